// K2 / K3 / K4: the per-step sampler updates as single vectorised (128-bit) element-wise
// kernels.  HBM-bound: each reads x and the UNet output once and writes x once.
//
//  K2  IADB   x' = (x + da*d[:, :C]) + dg*d[:, C:2C]   iadb_bn.py:326,329,344; utils.py:218-226;
//                                                       latent_iadb_bn_diffusers.py:110-117
//  K3  DDIM   x' = c2*clamp((x - c1*e)/c0) + c3*e [+ c4*noise]      diffusers DDIMScheduler.step
//                                                       (called at ddim_diffusers.py:680)
//  K4  u8     round(clamp(x/2+0.5,0,1)*255), NCHW -> NHWC            ddim_diffusers.py:687-688
//
// All arithmetic is explicit __f*_rn (no FMA contraction) in the reference's association so
// results are bit-identical to the torch expressions.
#include "common.cuh"

namespace bndm {

// Resident blocks per SM the step kernels are compiled for (<= 40 registers): the cfg-2 grid of
// 768 blocks must fit in ONE wave (measured: at 52 registers only 4 blocks fit, 1.3 waves, and
// the half-empty second wave doubled the kernel's duration).
constexpr int kStepBlocksPerSM = 6;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ float upd1(float x, float d, float a) { return __fadd_rn(x, __fmul_rn(a, d)); }
__device__ __forceinline__ float upd2(float x, float d1, float a, float d2, float g) {
  return __fadd_rn(__fadd_rn(x, __fmul_rn(a, d1)), __fmul_rn(g, d2));
}

// Scheduled steps keep their step index on the device so that one captured graph replays for
// every t: `state[0]` counts block tickets over the whole sampling run (zeroed by the host
// before the first step) and a launch of G blocks is step  ticket / G.  One atomic round trip
// per block, issued after the block's x / d loads so its latency hides behind them; nothing is
// written that a block of the same launch reads, so there is no end-of-kernel pass.  The block
// that draws the last ticket of a step publishes the next UNet timestep.
__device__ __forceinline__ int step_from_ticket(int *state, bool &publishes) {
  __shared__ int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&state[0], 1);
  __syncthreads();
  const int ticket = s_ticket;
  int step = ticket / (int)gridDim.x;
  publishes = (ticket - step * (int)gridDim.x) == (int)gridDim.x - 1;
  // state[1] = number of rows in the schedule table (0: unchecked).  A launch beyond the table (step_ called more
  // than T times without a reset, a replay loop that runs too long) re-uses the last row instead of reading past
  // the end, and raises the overrun flag the host can poll.
  const int n_rows = state[1] & 0x3FFFFFFF;
  if (n_rows > 0 && step >= n_rows) {
    step = n_rows - 1;
    if (threadIdx.x == 0 && publishes) atomicOr(&state[1], 0x40000000);
  }
  return step;
}

// table rows: [step][sample] x {dalpha, dgamma, t_next, 0} (per-sample, like the (B,) coefficient
// tensors the reference forms at iadb_bn.py:311-316)
template <bool kSched, bool kVec>
__global__ void __launch_bounds__(256, kStepBlocksPerSM) iadb_step_kernel(IadbArgs a) {
  const bool two = a.Cd == 2 * a.C;
  constexpr int V = kVec ? 4 : 1;
  const unsigned hwv = (unsigned)(a.HW / V);
  const unsigned total = (unsigned)a.B * a.C * hwv;          // launcher guarantees < 2^31
  const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;

  // software pipeline: an element's loads are issued one iteration ahead -- the first ones before
  // the ticket / schedule row (they do not depend on it), so those latencies overlap
  const unsigned stride = gridDim.x * blockDim.x;
  const int64_t plane = (int64_t)a.C * a.HW;
  unsigned idx = first;
  int b = 0;
  int64_t xo = 0;
  float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), u = xv, v = xv;
  auto fetch = [&]() {
    const unsigned bc = idx / hwv;
    const int hw = (int)(idx - bc * hwv) * V;
    b = (int)(bc / a.C);
    xo = (int64_t)bc * a.HW + hw;
    const int64_t d1 = xo + (int64_t)b * (a.Cd - a.C) * a.HW;       // ((b Cd + c) HW + hw)
    if (a.d_nhwc) {
      // channels-last UNet output d[b][hw][Cd]: channel c (and c + C) of V consecutive pixels
      const int c = (int)(bc - (unsigned)b * a.C);
      const float *dp = a.d + ((int64_t)b * a.HW + hw) * a.Cd + c;
      u.x = __ldg(dp);
      if (two) v.x = __ldg(dp + a.C);
      if (kVec) {
        u.y = __ldg(dp + a.Cd); u.z = __ldg(dp + 2 * a.Cd); u.w = __ldg(dp + 3 * a.Cd);
        if (two) { v.y = __ldg(dp + a.Cd + a.C); v.z = __ldg(dp + 2 * a.Cd + a.C); v.w = __ldg(dp + 3 * a.Cd + a.C); }
        xv = *reinterpret_cast<const float4 *>(a.x + xo);
      } else {
        xv.x = a.x[xo];
      }
    } else if (kVec) {
      xv = *reinterpret_cast<const float4 *>(a.x + xo);
      u = ldg4(a.d + d1);
      if (two) v = ldg4(a.d + d1 + plane);
    } else {
      xv.x = a.x[xo];
      u.x = __ldg(a.d + d1);
      if (two) v.x = __ldg(a.d + d1 + plane);
    }
  };
  if (idx < total) fetch();

  const float4 *rows = nullptr;
  bool publishes = false;
  if (kSched) {
    const int step = step_from_ticket(a.state, publishes);
    rows = reinterpret_cast<const float4 *>(a.table) + (int64_t)step * a.B;
  }

  while (idx < total) {
    float da, dg;
    if (kSched) {
      const float4 row = __ldg(rows + b);
      da = row.x; dg = row.y;
    } else {
      da = __ldg(a.dalpha + b);
      dg = two ? __ldg(a.dgamma + b) : 0.f;
    }
    float4 o;
    if (two) {
      o.x = upd2(xv.x, u.x, da, v.x, dg);
      if (kVec) { o.y = upd2(xv.y, u.y, da, v.y, dg); o.z = upd2(xv.z, u.z, da, v.z, dg); o.w = upd2(xv.w, u.w, da, v.w, dg); }
    } else {
      o.x = upd1(xv.x, u.x, da);
      if (kVec) { o.y = upd1(xv.y, u.y, da); o.z = upd1(xv.z, u.z, da); o.w = upd1(xv.w, u.w, da); }
    }
    float *dst = a.x_out + xo;
    idx += stride;
    if (idx < total) fetch();            // next element's loads go out before this store retires
    if (kVec) *reinterpret_cast<float4 *>(dst) = o;
    else *dst = o.x;
  }
  if (kSched && publishes && a.t_next_out)
    for (int b = threadIdx.x; b < a.B; b += blockDim.x) a.t_next_out[b] = __ldg(rows + b).z;
}

// Scheduled step with the UNet output in channels-last memory (d[b][hw][Cd]) and C <= 4: one thread per
// PIXEL -- consecutive lanes read consecutive 4*Cd-byte pixel records of d and consecutive floats of
// every x plane, so every access is coalesced (the generic kernel's per-channel mapping strides through
// d).  Same arithmetic, same ticket scheme.
constexpr int kNhwcMaxC = 4;
__global__ void __launch_bounds__(256, kStepBlocksPerSM) iadb_step_dnhwc_kernel(IadbArgs a) {
  const bool two = a.Cd == 2 * a.C;
  const unsigned total = (unsigned)a.B * a.HW;
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  float xv[kNhwcMaxC], u[kNhwcMaxC], v[kNhwcMaxC];
  int b = 0;
  unsigned p = 0;
  auto fetch = [&]() {
    b = (int)(idx / (unsigned)a.HW);
    p = idx - (unsigned)b * a.HW;
    const float *dp = a.d + (int64_t)idx * a.Cd;
#pragma unroll
    for (int c = 0; c < kNhwcMaxC; ++c)
      if (c < a.C) {
        xv[c] = a.x[((int64_t)b * a.C + c) * a.HW + p];
        u[c] = __ldg(dp + c);
        v[c] = two ? __ldg(dp + a.C + c) : 0.f;
      }
  };
  if (idx < total) fetch();
  bool publishes = false;
  const int step = step_from_ticket(a.state, publishes);
  const float4 *rows = reinterpret_cast<const float4 *>(a.table) + (int64_t)step * a.B;
  while (idx < total) {
    const float4 row = __ldg(rows + b);
#pragma unroll
    for (int c = 0; c < kNhwcMaxC; ++c)
      if (c < a.C)
        a.x_out[((int64_t)b * a.C + c) * a.HW + p] = two ? upd2(xv[c], u[c], row.x, v[c], row.y) : upd1(xv[c], u[c], row.x);
    idx += stride;                        // at most ~1.2 elements per thread: no look-ahead needed
    if (idx < total) fetch();
  }
  if (publishes && a.t_next_out)
    for (int bb = threadIdx.x; bb < a.B; bb += blockDim.x) a.t_next_out[bb] = __ldg(rows + bb).z;
}

static int grid_for(int64_t work_items, int threads) {
  // at most one full wave (148 SMs x kStepBlocksPerSM resident blocks); grid-stride covers the rest
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * kStepBlocksPerSM;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_iadb_step(const IadbArgs &a, bool sched, cudaStream_t s) {
  const bool vec = (a.HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.d) |
                                        reinterpret_cast<uintptr_t>(a.x_out)) % 16 == 0);
  const int64_t total = (int64_t)a.B * a.C * (a.HW / (vec ? 4 : 1));
  if (total >= (int64_t)1 << 31) return cudaErrorInvalidValue;      // 32-bit index math in the kernel
  if (sched && a.d_nhwc && a.C <= kNhwcMaxC) {
    iadb_step_dnhwc_kernel<<<grid_for((int64_t)a.B * a.HW, 256), 256, 0, s>>>(a);
    return cudaGetLastError();
  }
  const int grid = grid_for(total, 256);
  if (sched) {
    if (vec) iadb_step_kernel<true, true><<<grid, 256, 0, s>>>(a);
    else iadb_step_kernel<true, false><<<grid, 256, 0, s>>>(a);
  } else {
    if (vec) iadb_step_kernel<false, true><<<grid, 256, 0, s>>>(a);
    else iadb_step_kernel<false, false><<<grid, 256, 0, s>>>(a);
  }
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------- DDIM
__device__ __forceinline__ float ddim1(float x, float e, float z, bool has_noise, int clip, const float c[5]) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c[1], e)), c[0]);
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  float o = __fadd_rn(__fmul_rn(c[2], x0), __fmul_rn(c[3], e));
  if (has_noise) o = __fadd_rn(o, __fmul_rn(c[4], z));
  return o;
}

template <bool kVec>
__global__ void __launch_bounds__(256, kStepBlocksPerSM) ddim_step_kernel(DdimArgs a) {
  int step = 0;
  bool publishes = false;
  if (a.state) step = step_from_ticket(a.state, publishes);
  const float *row = a.coef + 8 * (int64_t)step;
  const float c[5] = {__ldg(row), __ldg(row + 1), __ldg(row + 2), __ldg(row + 3), __ldg(row + 4)};
  const float t_next = __ldg(row + 5);
  const bool hn = a.noise != nullptr;
  constexpr int V = kVec ? 4 : 1;
  const int64_t total = a.n / V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    if (kVec) {
      const float4 xv = *reinterpret_cast<const float4 *>(a.x + idx * 4);
      const float4 ev = ldg4(a.eps + idx * 4);
      float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hn) zv = ldg4(a.noise + idx * 4);
      float4 o;
      o.x = ddim1(xv.x, ev.x, zv.x, hn, a.clip, c); o.y = ddim1(xv.y, ev.y, zv.y, hn, a.clip, c);
      o.z = ddim1(xv.z, ev.z, zv.z, hn, a.clip, c); o.w = ddim1(xv.w, ev.w, zv.w, hn, a.clip, c);
      *reinterpret_cast<float4 *>(a.x_out + idx * 4) = o;
    } else {
      a.x_out[idx] = ddim1(a.x[idx], a.eps[idx], hn ? a.noise[idx] : 0.f, hn, a.clip, c);
    }
  }
  if (a.state && publishes && a.t_next_out)
    for (int b = threadIdx.x; b < a.B; b += blockDim.x) a.t_next_out[b] = t_next;
}

cudaError_t launch_ddim_step(const DdimArgs &a, cudaStream_t s) {
  uintptr_t al = reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.eps) | reinterpret_cast<uintptr_t>(a.x_out);
  if (a.noise) al |= reinterpret_cast<uintptr_t>(a.noise);
  const bool vec = (a.n % 4 == 0) && (al % 16 == 0);
  const int grid = grid_for(a.n / (vec ? 4 : 1), 256);
  if (vec) ddim_step_kernel<true><<<grid, 256, 0, s>>>(a);
  else ddim_step_kernel<false><<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------ uint8 NHWC
__global__ void __launch_bounds__(256) to_u8_kernel(const float *__restrict__ x, uint8_t *__restrict__ out, int B, int C,
                                                    int HW) {
  const int64_t total = (int64_t)B * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / HW, hw = idx - b * HW;
    for (int c = 0; c < C; ++c) {
      float v = __fadd_rn(__fdiv_rn(x[(b * C + c) * HW + hw], 2.0f), 0.5f);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      out[idx * C + c] = (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f));   // torch.round = half-to-even
    }
  }
}

// ---- the IADB test driver's PNG conversion (iadb_bn.py:796-816): one CTA per (C,H,W) image ------------------
//   final image:            v = clamp((x + 1) / 2, 0, 1)
//   intermediate snapshot:  v = (x - min(x)) / (max(x) - min(x))      (min / max over the whole image)
//   out[h][w][c] = (uint8)(v * 255)      numpy's astype(uint8): TRUNCATION (ddim_diffusers.py:687-688 rounds: K4 above)
// The image is read twice (the second pass hits L2); min / max by warp shuffles + one shared-memory round.
__global__ void __launch_bounds__(1024) snapshot_u8_kernel(const float *__restrict__ x, uint8_t *__restrict__ out, int C, int HW,
                                                           const int *__restrict__ final_flags, int final_all) {
  const int n = blockIdx.x;
  const float *xi = x + (int64_t)n * C * HW;
  const int total = C * HW;
  const bool fin = final_flags ? final_flags[n] != 0 : final_all != 0;
  float mn = 0.f, range = 1.f;
  if (!fin) {
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const float v = xi[i];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[32], s_hi[32];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    mn = lo;
    range = __fsub_rn(hi, lo);
  }
  uint8_t *oi = out + (int64_t)n * C * HW;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {        // i = hw * C + c (output order)
    const int hw = i / C, c = i - hw * C;
    const float v0 = xi[(int64_t)c * HW + hw];
    float v;
    if (fin) v = fminf(fmaxf(__fdiv_rn(__fadd_rn(v0, 1.0f), 2.0f), 0.0f), 1.0f);
    else v = __fdiv_rn(__fsub_rn(v0, mn), range);
    oi[i] = (uint8_t)__float2int_rz(__fmul_rn(v, 255.0f));
  }
}

cudaError_t launch_snapshot_u8(const float *x, uint8_t *out, int N, int C, int HW, const int *final_flags, int final_all,
                               cudaStream_t s) {
  snapshot_u8_kernel<<<N, 1024, 0, s>>>(x, out, C, HW, final_flags, final_all);
  return cudaGetLastError();
}

cudaError_t launch_to_u8(const float *x, uint8_t *out, int B, int C, int HW, cudaStream_t s) {
  to_u8_kernel<<<grid_for((int64_t)B * HW, 256), 256, 0, s>>>(x, out, B, C, HW);
  return cudaGetLastError();
}

}  // namespace bndm
