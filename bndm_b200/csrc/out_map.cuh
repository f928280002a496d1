// Output mapping shared by the combine kernel, the SIMT epilogue and the fused combine inside the
// tcgen05 contraction: everything get_noise_v2 does after the matmul (get_noise_recent.py:88-99,
// :113-118, :146-162) for GEMM column j / pixel p -- NCHW store, the white<->blue lerp with the
// per-sample gamma, the 32^2 crop, the 128^2 tile placement (noise_padding :7-19) and the 128^2
// noise_wn re-interpretation (:143-144).
#pragma once
#include "common.cuh"

namespace bndm {

struct OutMap {
  const float *z_cols;
  const float *gamma;
  float *out, *out_bn, *out_wn;
  int B, C, res_mode;
  TrainOut train = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// 4 consecutive pixels p..p+3 (same image row: p % 4 == 0) of GEMM column j, in two halves so the
// caller can put the white / gamma loads in flight together with its partial-tile loads.
struct OutPos4 {
  int64_t dst;     // flat output index of the first pixel, < 0: cropped away
  float4 wn;
  float g;
  float4 x1;       // training mode: the data pixels at dst
  float al, alp;   // training mode: alpha[b], alpha_prev[b]
};
__device__ __forceinline__ OutPos4 locate4(const OutMap &m, int j, int p) {
  const int h = p >> 6, w = p & 63;
  OutPos4 o;
  int b;
  o.wn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m.res_mode == kRes64) {
    b = j / m.C;
    o.dst = (int64_t)j * kNPix + p;
    o.wn = __ldg(reinterpret_cast<const float4 *>(m.z_cols + (int64_t)j * kNPix + p));
  } else if (m.res_mode == kRes32) {
    b = j / m.C;
    o.dst = (h >= 32 || w >= 32) ? -1 : (int64_t)j * 1024 + h * 32 + w;       // cropped away (:97-99)
    if (o.dst >= 0) o.wn = __ldg(reinterpret_cast<const float4 *>(m.z_cols + (int64_t)j * kNPix + p));
  } else {
    const int n = j / m.C, c = j - n * m.C;
    b = n >> 2;                           // (4B,...) re-viewed as (B,4,...): n = 4 b' + k'
    const int k = n & 3;
    const int r0 = (k & 1) * kTile, c0 = (k >> 1) * kTile;     // noise_padding placement
    o.dst = (((int64_t)b * m.C + c) * 128 + r0 + h) * 128 + c0 + w;
    // (n, pixel, channel) memory re-read as (n, channel, pixel): flat f = c*4096 + p
    const float *zn = m.z_cols + (int64_t)n * m.C * kNPix;
    float t[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = c * kNPix + p + e;
      const int cs = f % m.C, ps = f / m.C;
      t[e] = __ldg(zn + (int64_t)cs * kNPix + ps);
    }
    o.wn = make_float4(t[0], t[1], t[2], t[3]);
  }
  o.g = m.gamma ? __ldg(m.gamma + b) : 0.0f;
  o.x1 = make_float4(0.f, 0.f, 0.f, 0.f);
  o.al = o.alp = 0.0f;
  if (m.train.x_alpha && o.dst >= 0) {
    o.x1 = __ldg(reinterpret_cast<const float4 *>(m.train.x1 + o.dst));
    o.al = __ldg(m.train.alpha + b);
    o.alp = m.train.alpha_prev ? __ldg(m.train.alpha_prev + b) : 0.0f;
  }
  return o;
}
__device__ __forceinline__ void store4(const OutMap &m, const OutPos4 &q, float4 bn) {
  if (q.dst < 0) return;
  float4 o = bn;
  if (m.gamma) {
    const float g = q.g;
    const float gi = __fsub_rn(1.0f, g);
    o.x = __fadd_rn(__fmul_rn(bn.x, gi), __fmul_rn(q.wn.x, g));
    o.y = __fadd_rn(__fmul_rn(bn.y, gi), __fmul_rn(q.wn.y, g));
    o.z = __fadd_rn(__fmul_rn(bn.z, gi), __fmul_rn(q.wn.z, g));
    o.w = __fadd_rn(__fmul_rn(bn.w, gi), __fmul_rn(q.wn.w, g));
  }
  if (m.out) *reinterpret_cast<float4 *>(m.out + q.dst) = o;
  if (m.out_bn) *reinterpret_cast<float4 *>(m.out_bn + q.dst) = bn;
  if (m.out_wn) *reinterpret_cast<float4 *>(m.out_wn + q.dst) = q.wn;
  if (m.train.x_alpha) {
    const float a = q.al, ai = __fsub_rn(1.0f, q.al);
    float4 xa, t1;
    xa.x = __fadd_rn(__fmul_rn(a, o.x), __fmul_rn(ai, q.x1.x));
    xa.y = __fadd_rn(__fmul_rn(a, o.y), __fmul_rn(ai, q.x1.y));
    xa.z = __fadd_rn(__fmul_rn(a, o.z), __fmul_rn(ai, q.x1.z));
    xa.w = __fadd_rn(__fmul_rn(a, o.w), __fmul_rn(ai, q.x1.w));
    t1.x = __fsub_rn(q.x1.x, o.x); t1.y = __fsub_rn(q.x1.y, o.y); t1.z = __fsub_rn(q.x1.z, o.z); t1.w = __fsub_rn(q.x1.w, o.w);
    *reinterpret_cast<float4 *>(m.train.x_alpha + q.dst) = xa;
    *reinterpret_cast<float4 *>(m.train.tar1 + q.dst) = t1;
    if (m.train.tar2) {
      float4 t2;
      t2.x = __fmul_rn(q.alp, __fsub_rn(bn.x, q.wn.x)); t2.y = __fmul_rn(q.alp, __fsub_rn(bn.y, q.wn.y));
      t2.z = __fmul_rn(q.alp, __fsub_rn(bn.z, q.wn.z)); t2.w = __fmul_rn(q.alp, __fsub_rn(bn.w, q.wn.w));
      *reinterpret_cast<float4 *>(m.train.tar2 + q.dst) = t2;
    }
  }
}
__device__ __forceinline__ void emit4(const OutMap &m, int j, int p, float4 bn) { store4(m, locate4(m, j, p), bn); }

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// Scalar twin of emit4 in two halves, so a caller can issue the white-value loads of a batch of
// columns before any store (loads and stores may alias as far as the compiler knows, so a fused
// load+store per column would serialise one DRAM round trip per column -- measured).
struct OutPos {
  int64_t dst;     // flat output index, < 0: pixel is cropped away
  int b;           // sample (for gamma)
};
__device__ __forceinline__ OutPos locate1(const OutMap &m, int j, int p, float &wn) {
  const int h = p >> 6, w = p & 63;
  OutPos o;
  if (m.res_mode == kRes64) {
    o.b = j / m.C;
    o.dst = (int64_t)j * kNPix + p;
    wn = __ldg(m.z_cols + (int64_t)j * kNPix + p);
  } else if (m.res_mode == kRes32) {
    o.b = j / m.C;
    o.dst = (h >= 32 || w >= 32) ? -1 : (int64_t)j * 1024 + h * 32 + w;       // cropped away (:97-99)
    wn = __ldg(m.z_cols + (int64_t)j * kNPix + p);
  } else {
    const int n = j / m.C, c = j - n * m.C;
    o.b = n >> 2;                          // (4B,...) re-viewed as (B,4,...): n = 4 b' + k'
    const int k = n & 3;
    const int r0 = (k & 1) * kTile, c0 = (k >> 1) * kTile;      // noise_padding placement
    o.dst = (((int64_t)o.b * m.C + c) * 128 + r0 + h) * 128 + c0 + w;
    const int f = c * kNPix + p;           // (n, pixel, channel) memory re-read as (n, channel, pixel)
    const int cs = f % m.C, ps = f / m.C;
    wn = __ldg(m.z_cols + ((int64_t)n * m.C + cs) * kNPix + ps);
  }
  return o;
}
__device__ __forceinline__ void store1(const OutMap &m, const OutPos &o, float bn, float wn, float g) {
  if (o.dst < 0) return;
  float v = bn;
  if (m.gamma) v = __fadd_rn(__fmul_rn(bn, __fsub_rn(1.0f, g)), __fmul_rn(wn, g));
  m.out[o.dst] = v;
  if (m.out_bn) m.out_bn[o.dst] = bn;
  if (m.out_wn) m.out_wn[o.dst] = wn;
}
// N columns col .. col + N - 1 (those < n_cols) of pixel p: all loads, then all stores
template <int N>
__device__ __forceinline__ void emit_cols(const OutMap &m, int col, int n_cols, int p, const float *bn) {
  OutPos pos[N];
  float wn[N], g[N];
#pragma unroll
  for (int e = 0; e < N; ++e) {
    wn[e] = 0.0f;
    g[e] = 0.0f;
    pos[e].dst = -1;
    if (col + e < n_cols) {
      pos[e] = locate1(m, col + e, p, wn[e]);
      if (m.gamma) g[e] = __ldg(m.gamma + pos[e].b);
    }
  }
#pragma unroll
  for (int e = 0; e < N; ++e) store1(m, pos[e], bn[e], wn[e], g[e]);
}

}  // namespace bndm
