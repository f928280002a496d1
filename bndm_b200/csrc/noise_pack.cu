// K1a "pack": gathers the white field into GEMM-column order, splits it for 3xTF32, and
// the small helper kernels around it (white re-interpretation, L split, triangularity check).
//
// Column order: j = n*C + c, n = sample (64^2, 32^2) or 64x64 tile (128^2: n = k*B + b when
// the source is the caller's image, get_noise_recent.py:131-132; the draw order when the
// source is a (4B,C,64,64) draw, :138).  Row p = h*64 + w inside the tile (:111).
#include <stdlib.h>

#include "common.cuh"

namespace bndm {

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

__global__ void __launch_bounds__(256) pack_kernel(PackArgs a) {
  // one thread per float4 of one packed column: 1024 float4 per column
  pdl_launch_dependents();      // the contraction may start its set-up and L loads now
  pdl_wait();                   // the previous call's kernels are done with zt / z_raw
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
      // tile n of the GLOBAL dim-0 concatenation (a shard owns tiles [n0, n0 + 4 B) of the 4 Bg)
      const int nl = j / a.C, c = j - nl * a.C;
      const int n = nl + a.n0;
      const int k = n / a.Bg, b = n - k * a.Bg;
      const int r0 = (k >> 1) * kTile, c0 = (k & 1) * kTile;
      v = ld4(a.src + (((int64_t)b * a.C + c) * 128 + r0 + h) * 128 + c0 + w);
    }
  }
  const int64_t o = (int64_t)j * kNPix + p;
  if (a.z_raw) st4(a.z_raw + o, v);
  if (a.zt) {
    float4 hi, lo;
    tf32_split(v.x, hi.x, lo.x);
    tf32_split(v.y, hi.y, lo.y);
    tf32_split(v.z, hi.z, lo.z);
    tf32_split(v.w, hi.w, lo.w);
    // stage block (cb, s): rows 0..nb-1 = zh of the block's columns, rows nb..2nb-1 = zl
    const int cb = j / a.nb, jc = j - cb * a.nb;
    const int s = p >> 5, kk = p & 31;
    uint8_t *blk = reinterpret_cast<uint8_t *>(a.zt) + ((size_t)cb * (kNPix / kStageK) + s) * ((size_t)2 * a.nb * kStageK * 4);
    *reinterpret_cast<float4 *>(blk + sw128_offset(jc, kk)) = hi;
    *reinterpret_cast<float4 *>(blk + sw128_offset(a.nb + jc, kk)) = lo;
  }
}

cudaError_t launch_pack(const PackArgs &a, cudaStream_t s) {
  const int64_t n4 = (int64_t)a.n_cols_pad * 1024;
  return launch_pdl(pack_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, a);
}

// ---- 'gaussian' 128^2 test-mode pass-through (get_noise_recent.py:50-56) ------------------
// out[b', c', r0+h, c0+w] = x_tile[n][f % C][f / C],  n = 4 b' + k', f = c'*4096 + h*64 + w,
// x_tile[n = k*Bg + b] = quadrant k of image b of the GLOBAL batch Bg (the shard [b0, b0+B) computes its own
// output images only); placement (r0,c0) = ((k'&1)*64, (k'>>1)*64).
__global__ void __launch_bounds__(256) white128_kernel(const float *__restrict__ x, float *__restrict__ out,
                                                       int B, int C, int Bg, int b0) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*C*128*128 (this shard's outputs)
  const int64_t total = (int64_t)B * C * 128 * 128;
  if (idx >= total) return;
  const int X = (int)(idx & 127), Y = (int)((idx >> 7) & 127);
  const int bc = (int)(idx >> 14);
  const int bo = bc / C, co = bc - bo * C;
  const int kq = ((X >> 6) << 1) | (Y >> 6);          // inverse of the output placement
  const int h = Y & 63, w = X & 63;
  const int n = (bo + b0) * 4 + kq;                   // tile index in the GLOBAL (4 Bg, ...) concatenation
  const int f = co * kNPix + h * kTile + w;
  const int c = f % C, p = f / C;
  const int k = n / Bg, b = n - k * Bg;
  const int r0 = (k >> 1) * kTile, c0 = (k & 1) * kTile;
  out[idx] = x[(((int64_t)b * C + c) * 128 + r0 + (p >> 6)) * 128 + c0 + (p & 63)];
}

cudaError_t launch_white128(const float *x, float *out, int B, int C, int Bg, int b0, cudaStream_t s) {
  const int64_t total = (int64_t)B * C * 128 * 128;
  white128_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, out, B, C, Bg, b0);
  return cudaGetLastError();
}

// ---- init-time helpers ---------------------------------------------------------------------
// mode 0: hi = rna(v), lo = rna(v - hi) (production).  Experiment modes (BNDM_L_SPLIT_MODE, which
// probe how tcgen05 kind::tf32 reads a full fp32 operand): 1: hi = v as is, lo = v - trunc(v);
// 2: hi = v as is, lo = v - rna(v); 3: hi = trunc(v), lo = v - hi (explicit truncation split).
__device__ __forceinline__ void split_mode(float v, int mode, float &hi, float &lo) {
  if (mode == 0) { tf32_split(v, hi, lo); return; }
  const float tr = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  if (mode == 1) { hi = v; lo = __fsub_rn(v, tr); }
  else if (mode == 2) { hi = v; lo = __fsub_rn(v, tf32_rna(v)); }
  else { hi = tr; lo = __fsub_rn(v, tr); }
}

// one CTA per stage block: 128 rows x 32 k of L -> [Lh tile | Ll tile], SWIZZLE_128B image
// raw != 0: 16 KiB blocks holding the fp32 tile itself (converter variant of the contraction)
__global__ void __launch_bounds__(256) tile_L_kernel(const float *__restrict__ L, float *__restrict__ Lt, int dense, int mode,
                                                     int raw) {
  const int blk = blockIdx.x;
  int tile, s;
  if (dense) {
    tile = blk >> 7;
    s = blk & 127;
  } else {
    tile = 0;
    while (2 * (tile + 1) * (tile + 2) <= blk) ++tile;
    s = blk - 2 * tile * (tile + 1);
  }
  uint8_t *dst = reinterpret_cast<uint8_t *>(Lt + (size_t)blk * (raw ? kLBlockFloats / 2 : kLBlockFloats));
  for (int f = threadIdx.x; f < kBlk * (kStageK / 4); f += blockDim.x) {
    const int r = f >> 3, kk = (f & 7) << 2;
    const float4 v = ld4(L + (size_t)(tile * kBlk + r) * kNPix + s * kStageK + kk);
    if (raw) {
      *reinterpret_cast<float4 *>(dst + sw128_offset(r, kk)) = v;
      continue;
    }
    float4 a, b;
    split_mode(v.x, mode, a.x, b.x);
    split_mode(v.y, mode, a.y, b.y);
    split_mode(v.z, mode, a.z, b.z);
    split_mode(v.w, mode, a.w, b.w);
    const uint32_t o = sw128_offset(r, kk);
    *reinterpret_cast<float4 *>(dst + o) = a;
    *reinterpret_cast<float4 *>(dst + kBlk * kStageK * 4 + o) = b;
  }
}

cudaError_t launch_tile_L(const float *L, float *Lt, int dense, int raw, cudaStream_t s) {
  int mode = 0;
  if (const char *e = getenv("BNDM_L_SPLIT_MODE")) mode = atoi(e);
  tile_L_kernel<<<(unsigned)tile_L_blocks(dense), 256, 0, s>>>(L, Lt, dense, mode, raw);
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
