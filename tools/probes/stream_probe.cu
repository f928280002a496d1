// Stand-alone probes for the questions the round-1 experiment log leaves open about get_noise at small N
// (profiles/r01_experiment_log.md, "Round-2 work plan").  Not part of libbndm_b200.so.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/stream_probe tools/probes/stream_probe.cu
//   build/stream_probe [reps=7] [clean=0]     # prints one line per configuration; clean=1 flushes L2 with reads
//
// (1) tma   : G persistent CTAs each stream a contiguous slice of a 33.5 MB buffer (= the lower triangle of L)
//             with linear cp.async.bulk requests of `req` bytes, `depth` in flight, data not consumed: the
//             ceiling for K1b's streaming phase and the cold time-to-first-data, from a flushed L2.
// (2) ldg   : the same slice with 128-bit ld.global.nc from all threads (what a SIMT kernel would see).
// (3) fence : after streaming, every CTA stores an 8 KiB partial tile, __threadfence()s and bumps a counter
//             (the hand-off the in-kernel combine needs): how long does the fence take while others still stream?
// (4) hold  : (1) with a consumer that keeps each landed stage for 300 / 600 ns of serial work, at several ring depths
//             and with cp.async.bulk.prefetch.L2 look-ahead: isolates why deeper rings / look-ahead lost inside K1b.
// Times are %globaltimer stamps taken in the kernel: span = max(end) - min(start) over CTAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e__ = (x);                                                                        \
    if (e__ != cudaSuccess) {                                                                     \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__));         \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

__device__ __forceinline__ uint64_t gtime() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();        // bounded: a protocol bug traps instead of hanging the box
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Stamps {
  uint64_t start, first, end, fence_begin, fence_end;
};

constexpr int kMaxDepth = 8;

// (1) + (3): thread 0 is producer and consumer at once: it keeps `depth` requests in flight and re-arms a slot
// as soon as its bytes have landed.  Nothing reads the data: this is the TMA/DRAM ceiling.
// hold_ns > 0 emulates a consumer that keeps a landed stage for that long before the slot is re-armed (K1b: ~1200 ns
// for conversion + MMA issue + completion); ahead > 0 keeps that many requests beyond the ring prefetched into L2.
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(128) tma_stream_kernel(const char *src, size_t bytes_per_cta, uint32_t req, int depth,
                                                         float *partials, unsigned *counter, int do_fence, int hold_ns,
                                                         int ahead, Stamps *st) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t full[kMaxDepth];
  const char *mine = src + (size_t)blockIdx.x * bytes_per_cta;
  const int n_req = (int)(bytes_per_cta / req);
  uint64_t t0 = 0, t_first = 0;
  if (threadIdx.x == 0) {
    t0 = gtime();
    for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < depth && i < n_req; ++i) {
      mbar_expect_tx(&full[i], req);
      bulk_load(smem_u32(smem + (size_t)i * req), mine + (size_t)i * req, req, &full[i]);
    }
    for (int i = depth; i < depth + ahead && i < n_req; ++i) bulk_prefetch_l2(mine + (size_t)i * req, req);
    uint64_t busy_until = 0;                       // the emulated consumer handles one stage at a time
    for (int i = 0; i < n_req; ++i) {
      const int slot = i % depth;
      mbar_wait(&full[slot], (i / depth) & 1);
      if (i == 0) t_first = gtime();
      if (hold_ns > 0) {
        const uint64_t now = gtime();
        busy_until = (now > busy_until ? now : busy_until) + (uint64_t)hold_ns;
        while (gtime() < busy_until) {}
      }
      const int nxt = i + depth;
      if (ahead > 0 && nxt + ahead < n_req) bulk_prefetch_l2(mine + (size_t)(nxt + ahead) * req, req);
      if (nxt < n_req) {
        mbar_expect_tx(&full[slot], req);
        bulk_load(smem_u32(smem + (size_t)slot * req), mine + (size_t)nxt * req, req, &full[slot]);
      }
    }
  }
  __syncthreads();
  uint64_t f0 = 0, f1 = 0;
  if (do_fence) {                                                   // (3) the hand-off of an 8 KiB partial tile
    f0 = gtime();
    float4 *p = reinterpret_cast<float4 *>(partials + (size_t)blockIdx.x * 2048);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) p[i] = make_float4(1.f, 2.f, 3.f, (float)i);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(counter, 1u);
    f1 = gtime();
  }
  if (threadIdx.x == 0) st[blockIdx.x] = Stamps{t0, t_first, gtime(), f0, f1};
}

// (2) all threads read 128-bit words, 4 independent loads in flight per thread
__global__ void __launch_bounds__(512) ldg_stream_kernel(const float4 *src, size_t vec_per_cta, float *sink, Stamps *st) {
  const float4 *mine = src + (size_t)blockIdx.x * vec_per_cta;
  uint64_t t0 = gtime(), t_first = 0;
  float acc = 0.f;
  size_t i = threadIdx.x;
  const size_t stride = blockDim.x;
  for (; i + 3 * stride < vec_per_cta; i += 4 * stride) {
    float4 a = __ldg(mine + i), b = __ldg(mine + i + stride), c = __ldg(mine + i + 2 * stride), d = __ldg(mine + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
    if (t_first == 0) t_first = gtime();
  }
  for (; i < vec_per_cta; i += stride) acc += __ldg(mine + i).x;
  if (acc == 12345.678f) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0) st[blockIdx.x] = Stamps{t0, t_first, gtime(), 0, 0};
}

__global__ void flush_kernel(float4 *buf, size_t n, int clean, float *sink) {
  // clean == 0: dirties > L2 worth of lines (their write-backs then share the DRAM bus with the probe's reads);
  // clean == 1: reads them instead, leaving clean lines that are dropped without traffic
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (clean) acc += buf[i].x;
    else buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (clean && acc == 12345.678f) sink[1] = acc;
}

struct Result {
  double span_us, first_us_mean, first_us_min, fence_us_mean, fence_us_max;
};

static Result reduce(const std::vector<Stamps> &s, bool fence) {
  uint64_t lo = ~0ull, hi = 0;
  double fsum = 0, fmin = 1e30, qsum = 0, qmax = 0;
  for (const Stamps &x : s) {
    lo = std::min(lo, x.start);
    hi = std::max(hi, x.end);
    double f = (double)(x.first - x.start) * 1e-3;
    fsum += f;
    fmin = std::min(fmin, f);
    if (fence) {
      double q = (double)(x.fence_end - x.fence_begin) * 1e-3;
      qsum += q;
      qmax = std::max(qmax, q);
    }
  }
  return Result{(double)(hi - lo) * 1e-3, fsum / s.size(), fmin, qsum / s.size(), qmax};
}

int main(int argc, char **argv) {
  const size_t total = 33562624 / (148 * 65536) * (size_t)(148 * 65536) + 148 * 65536;     // ~33.5 MB, divisible for every config
  const int reps = argc > 1 ? atoi(argv[1]) : 7;
  const int clean = argc > 2 ? atoi(argv[2]) : 0;                    // 1: flush L2 with reads (clean lines)
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("# device %s, %d SMs, buffer %.2f MB, %d cold repetitions (median), L2 flushed with %s\n", prop.name, sms, total * 1e-6, reps,
         clean ? "reads (clean lines)" : "writes (dirty lines)");
  char *src;
  float4 *flush;
  float *partials, *sink;
  unsigned *counter;
  Stamps *st_dev;
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&src, total + (1 << 20)));
  CK(cudaMemset(src, 1, total + (1 << 20)));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMalloc(&partials, 4 * sms * 2048 * sizeof(float)));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMalloc(&counter, 16));
  CK(cudaMemset(counter, 0, 16));
  CK(cudaMalloc(&st_dev, 4 * sms * sizeof(Stamps)));
  CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

  auto run = [&](const char *name, int G, uint32_t req, int depth, int fence, int threads, int hold_ns = 0, int ahead = 0) {
    std::vector<double> span, first, firstmin, fmean, fmax;
    std::vector<Stamps> st(G);
    size_t per = total / G;
    if (req) per = per / req * req;
    for (int r = 0; r < reps + 1; ++r) {
      flush_kernel<<<sms * 4, 512>>>(flush, flush_bytes / 16, clean, sink);
      if (req)
        tma_stream_kernel<<<G, 128, (size_t)req * depth>>>(src, per, req, depth, partials, counter, fence, hold_ns, ahead, st_dev);
      else
        ldg_stream_kernel<<<G, threads>>>(reinterpret_cast<const float4 *>(src), per / 16, sink, st_dev);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(st.data(), st_dev, G * sizeof(Stamps), cudaMemcpyDeviceToHost));
      if (r == 0) continue;                                          // first launch: module load
      Result x = reduce(st, fence != 0);
      span.push_back(x.span_us);
      first.push_back(x.first_us_mean);
      firstmin.push_back(x.first_us_min);
      fmean.push_back(x.fence_us_mean);
      fmax.push_back(x.fence_us_max);
    }
    auto med = [](std::vector<double> v) {
      std::sort(v.begin(), v.end());
      return v[v.size() / 2];
    };
    const double bytes = (double)per * G;
    printf("%-5s G=%3d req=%3uK depth=%d thr=%4d | span %6.2f us  %5.2f TB/s | first data mean %5.2f min %5.2f us", name, G,
           req >> 10, depth, threads, med(span), bytes / med(span) * 1e-6, med(first), med(firstmin));
    if (fence) printf(" | store+fence+atomic mean %5.2f max %5.2f us", med(fmean), med(fmax));
    if (hold_ns || ahead) printf(" | consumer holds a stage %d ns, L2 look-ahead %d requests", hold_ns, ahead);
    printf("\n");
  };

  for (uint32_t req : {8u << 10, 16u << 10, 32u << 10, 64u << 10})
    for (int depth : {2, 3, 4, 6})
      if ((size_t)req * depth <= 192 * 1024) run("tma", sms, req, depth, 0, 128);
  run("tma", 2 * sms, 16u << 10, 3, 0, 128);
  run("tma", 2 * sms, 32u << 10, 3, 0, 128);
  run("tma", sms, 32u << 10, 3, 1, 128);
  run("tma", sms, 32u << 10, 4, 1, 128);
  // a consumer that holds every stage (K1b: conversion + MMA issue + completion ~ 600 ns serial work per 32 KiB stage,
  // the slot is free ~1200 ns after landing): how much ring depth / L2 look-ahead does the stream need then?
  for (int hold : {300, 600})
    for (int depth : {3, 4, 5, 6}) run("tma", sms, 32u << 10, depth, 0, 128, hold, 0);
  for (int ahead : {2, 4, 8}) run("tma", sms, 32u << 10, 3, 0, 128, 600, ahead);
  for (int threads : {256, 512})
    for (int mult : {1, 2, 4}) run("ldg", mult * sms, 0, 0, 0, threads);
  return 0;
}
