// pipe_probe: issue / throughput facts the K1g design hinges on, measured with clock64 inside one CTA per SM
// (16 warps = 4 per SM sub-partition, like K1g's consumers).  `make probes && build/pipe_probe` on a B200.
//   ffma    : scalar FFMA, 48 independent accumulators per thread (register-blocked inner product)
//   ffma2   : fma.rn.f32x2, 24 accumulator pairs
//   mma     : mma.sync.m16n8k8 tf32, 8 independent accumulator tiles per warp
//   lds     : LDS.128 with (a) one address per warp, (b) one 128-byte segment per quarter-warp (4 identical segments),
//             (c) 512 contiguous bytes per warp, (d) LDS.32 one address per warp
// Prints cycles per warp-instruction per sub-partition (ffma / ffma2 / mma) or per SM (lds) and the implied rates.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kWarps = 16, kIters = 512;

__global__ void __launch_bounds__(kWarps * 32, 1) ffma_kernel(float *out, long long *cyc, float a0, float b0) {
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = (float)i;
  float a[4] = {a0, a0 + 1.f, a0 + 2.f, a0 + 3.f};
  float b[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) b[j] = b0 + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 12; ++j) acc[i * 12 + j] = fmaf(a[i], b[j], acc[i * 12 + j]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 48; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void fma2(float2 &d, const float2 &a, const float2 &b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
               : "+l"(reinterpret_cast<unsigned long long &>(d))
               : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

__global__ void __launch_bounds__(kWarps * 32, 1) ffma2_kernel(float *out, long long *cyc, float a0, float b0) {
  float2 acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = make_float2((float)i, (float)-i);
  float2 a[2] = {make_float2(a0, a0 + 1.f), make_float2(a0 + 2.f, a0 + 3.f)};
  float2 b[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) b[j] = make_float2(b0 + j, b0 - j);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 12; ++j) fma2(acc[i * 12 + j], a[i], b[j]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int TILES>
__global__ void __launch_bounds__(kWarps * 32, 1) mma_kernel(float *out, long long *cyc, uint32_t a0, uint32_t b0) {
  float d[TILES][4];
#pragma unroll
  for (int i = 0; i < TILES; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  uint32_t a[2][4], b[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) a[i][e] = a0 + 0x1000u * (i * 4 + e) + threadIdx.x;
    b[i][0] = b0 + 0x2000u * i;
    b[i][1] = b0 + 0x4000u * i + threadIdx.x;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < TILES; ++i) mma_tf32(d[i], a[i & 1], b[(i >> 1) & 1]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TILES; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 1) lds_kernel(float *out, long long *cyc) {
  __shared__ __align__(16) float buf[8 * 1024];
  for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) buf[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off;        // float index
  if (MODE == 0 || MODE == 3) off = warp * 512;
  else if (MODE == 1) off = warp * 512 + (lane & 7) * 4;
  else off = warp * 512 + lane * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int o = (off + u * 128 + (it & 3) * 1024) & (8 * 1024 - 1);
      if (MODE == 3) {
        s.x += buf[o];
      } else {
        const float4 v = *reinterpret_cast<const float4 *>(buf + o);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s.x + s.y + s.z + s.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double mean_cyc(long long *h, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)h[i];
  return s / n;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  float *out;
  long long *cyc, *h = new long long[sms];
  CK(cudaMalloc(&out, (size_t)sms * kWarps * 32 * sizeof(float)));
  CK(cudaMalloc(&cyc, sms * sizeof(long long)));
  printf("pipe_probe: %d SMs, max clock %d MHz, %d warps per CTA (4 per sub-partition), %d iterations\n", sms, khz / 1000, kWarps, kIters);
  const int wps = kWarps / 4;
  for (int rep = 0; rep < 2; ++rep) {
    ffma_kernel<<<sms, kWarps * 32>>>(out, cyc, 1.0f, 2.0f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double c = mean_cyc(h, sms);
    if (rep) printf("ffma   : %.2f cycles per warp-FFMA per sub-partition  (%.1f FMA/clk/SM)\n", c / (kIters * 48.0 * wps), 128.0 / (c / (kIters * 48.0 * wps)));
    ffma2_kernel<<<sms, kWarps * 32>>>(out, cyc, 1.0f, 2.0f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    c = mean_cyc(h, sms);
    if (rep) printf("ffma2  : %.2f cycles per warp-FFMA2 per sub-partition (%.1f FMA/clk/SM)\n", c / (kIters * 24.0 * wps), 256.0 / (c / (kIters * 24.0 * wps)));
    mma_kernel<8><<<sms, kWarps * 32>>>(out, cyc, 0x3f800000u, 0x3f000000u);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    c = mean_cyc(h, sms);
    if (rep) printf("mma x8 : %.2f cycles per mma.sync.m16n8k8.tf32 per sub-partition (%.0f MAC/clk/SM)\n", c / (kIters * 8.0 * wps), 4 * 1024.0 / (c / (kIters * 8.0 * wps)));
    mma_kernel<2><<<sms, kWarps * 32>>>(out, cyc, 0x3f800000u, 0x3f000000u);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    c = mean_cyc(h, sms);
    if (rep) printf("mma x2 : %.2f cycles per mma.sync (2 accumulator tiles per warp: latency-bound check)\n", c / (kIters * 2.0 * wps));
    const char *names[4] = {"LDS.128, one address per warp", "LDS.128, one 128-byte segment per quarter-warp (x4 identical)",
                            "LDS.128, 512 contiguous bytes per warp", "LDS.32, one address per warp"};
    for (int m = 0; m < 4; ++m) {
      if (m == 0) lds_kernel<0><<<sms, kWarps * 32>>>(out, cyc);
      if (m == 1) lds_kernel<1><<<sms, kWarps * 32>>>(out, cyc);
      if (m == 2) lds_kernel<2><<<sms, kWarps * 32>>>(out, cyc);
      if (m == 3) lds_kernel<3><<<sms, kWarps * 32>>>(out, cyc);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
      c = mean_cyc(h, sms);
      if (rep) printf("lds %d  : %.2f cycles per warp-load per SM   (%s)\n", m, c / (kIters * 8.0 * kWarps), names[m]);
    }
  }
  return 0;
}
