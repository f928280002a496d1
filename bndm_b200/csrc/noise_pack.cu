// K1a "pack": gathers the white field into GEMM-column order, splits it for 3xTF32, and
// the small helper kernels around it (white re-interpretation, L split, triangularity check).
//
// Column order: j = n*C + c, n = sample (64^2, 32^2) or 64x64 tile (128^2: n = k*B + b when
// the source is the caller's image, get_noise_recent.py:131-132; the draw order when the
// source is a (4B,C,64,64) draw, :138).  Row p = h*64 + w inside the tile (:111).
#include "common.cuh"

namespace bndm {

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

__global__ void __launch_bounds__(256) pack_kernel(PackArgs a) {
  // one thread per float4 of one packed column: 1024 float4 per column
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)(idx >> 10);
  if (j >= a.n_cols_pad) return;
  const int p = ((int)idx & 1023) << 2;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < a.n_cols) {
    const int h = p >> 6, w = p & 63;
    if (!a.src_is_image || a.res_mode == kRes64) {
      v = ld4(a.src + (int64_t)j * kNPix + p);
    } else if (a.res_mode == kRes32) {          // 2x2 periodic tiling of a 32x32 image (:78-79)
      v = ld4(a.src + (int64_t)j * 1024 + (h & 31) * 32 + (w & 31));
    } else {                                    // quadrant k of image b (:131-132)
      const int n = j / a.C, c = j - n * a.C;
      const int k = n / a.B, b = n - k * a.B;
      const int r0 = (k >> 1) * kTile, c0 = (k & 1) * kTile;
      v = ld4(a.src + (((int64_t)b * a.C + c) * 128 + r0 + h) * 128 + c0 + w);
    }
  }
  const int64_t o = (int64_t)j * kNPix + p;
  if (a.z_raw) st4(a.z_raw + o, v);
  if (a.z_hi) {
    float4 hi, lo;
    tf32_split(v.x, hi.x, lo.x);
    tf32_split(v.y, hi.y, lo.y);
    tf32_split(v.z, hi.z, lo.z);
    tf32_split(v.w, hi.w, lo.w);
    st4(a.z_hi + o, hi);
    st4(a.z_lo + o, lo);
  }
}

cudaError_t launch_pack(const PackArgs &a, cudaStream_t s) {
  const int64_t n4 = (int64_t)a.n_cols_pad * 1024;
  pack_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ---- 'gaussian' 128^2 test-mode pass-through (get_noise_recent.py:50-56) ------------------
// out[b', c', r0+h, c0+w] = x_tile[n][f % C][f / C],  n = 4 b' + k', f = c'*4096 + h*64 + w,
// x_tile[n = k*B + b] = quadrant k of image b; placement (r0,c0) = ((k'&1)*64, (k'>>1)*64).
__global__ void __launch_bounds__(256) white128_kernel(const float *__restrict__ x, float *__restrict__ out,
                                                       int B, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*C*128*128
  const int64_t total = (int64_t)B * C * 128 * 128;
  if (idx >= total) return;
  const int X = (int)(idx & 127), Y = (int)((idx >> 7) & 127);
  const int bc = (int)(idx >> 14);
  const int bo = bc / C, co = bc - bo * C;
  const int kq = ((X >> 6) << 1) | (Y >> 6);          // inverse of the output placement
  const int h = Y & 63, w = X & 63;
  const int n = bo * 4 + kq;
  const int f = co * kNPix + h * kTile + w;
  const int c = f % C, p = f / C;
  const int k = n / B, b = n - k * B;
  const int r0 = (k >> 1) * kTile, c0 = (k & 1) * kTile;
  out[idx] = x[(((int64_t)b * C + c) * 128 + r0 + (p >> 6)) * 128 + c0 + (p & 63)];
}

cudaError_t launch_white128(const float *x, float *out, int B, int C, cudaStream_t s) {
  const int64_t total = (int64_t)B * C * 128 * 128;
  white128_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, out, B, C);
  return cudaGetLastError();
}

// ---- init-time helpers ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ src, float *__restrict__ hi,
                                                         float *__restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld4(src + i * 4);
    float4 a, b;
    tf32_split(v.x, a.x, b.x);
    tf32_split(v.y, a.y, b.y);
    tf32_split(v.z, a.z, b.z);
    tf32_split(v.w, a.w, b.w);
    st4(hi + i * 4, a);
    st4(lo + i * 4, b);
  }
}

cudaError_t launch_split_tf32(const float *src, float *hi, float *lo, int64_t n, cudaStream_t s) {
  split_tf32_kernel<<<148 * 8, 256, 0, s>>>(src, hi, lo, n / 4);
  return cudaGetLastError();
}

// flag = 1 if any element strictly above the diagonal is non-zero
__global__ void __launch_bounds__(256) tri_check_kernel(const float *__restrict__ L, int n, int *flag) {
  const int64_t total = (int64_t)n * n;
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n), c = (int)(i - (int64_t)r * n);
    if (c > r && L[i] != 0.0f) bad = 1;
  }
  if (bad) atomicOr(flag, 1);
}

cudaError_t launch_tri_check(const float *L, int n, int *flag_dev, cudaStream_t s) {
  tri_check_kernel<<<148 * 8, 256, 0, s>>>(L, n, flag_dev);
  return cudaGetLastError();
}

}  // namespace bndm
