// K5: fused [residual add] + [per-sample channel bias] + GroupNorm + [SiLU] on channels-last
// (NHWC) fp32 activations -- the glue between the UNet's convolutions (diffusers ResnetBlock2D:
// norm1/act/conv1, "+ time_emb_proj(temb)[:, :, None, None]", norm2/act/conv2, "+ shortcut";
// the UNet the reference samples with, iadb_bn.py:205-282 / :319).
//
//     s = x (+ res) (+ add_bc[b, c])          optionally written to sum_out (the next residual)
//     y = act( (s - mean_g) * rstd_g * weight[c] + bias[c] )        act = SiLU or identity
//
// PyTorch runs this as RowwiseMoments + an element-wise affine kernel + a SiLU kernel (+ separate
// add kernels): 3-4 reads and 2-3 writes of the activation; here it is ONE kernel that reads the
// activation from HBM once (the second, normalising pass re-reads it from L2) and writes y once.
//
// One CTA per (sample, block of whole groups spanning a multiple of 32 channels): every pixel
// contributes one or three full 128-byte lines, so all global traffic is coalesced without any
// transposition.  A thread owns one float4 channel quad (always inside one group since
// channels-per-group is a multiple of 4) and walks the pixels; statistics are per-thread shifted
// sums merged with Chan's parallel-variance formula in a FIXED order (bit-reproducible).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace bndm {

struct GnArgs {
  const float *x, *res, *add_bc, *weight, *bias;
  const float *x2;         // second source (cluster kernel): channels [C1, C) come from x2 ([B][HW][C - C1]); null = single source
  int C1;                  // channels held by x (== C when x2 is null)
  float *sum_out, *y;
  int B, C, HW, cpg;       // cpg = channels per group (multiple of 4)
  int add_stride;          // floats between consecutive samples' rows of add_bc (>= C)
  int cblk;                // channels per CTA (multiple of 32 and of cpg)
  float eps;
  int silu;
};

struct Moments {
  float n, mean, m2;
};
__device__ __forceinline__ Moments merge(const Moments &a, const Moments &b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float n = a.n + b.n;
  const float d = b.mean - a.mean;
  Moments r;
  r.n = n;
  const float f = __fdividef(b.n, n);             // SFU reciprocal: deterministic, ~2 ulp
  r.mean = a.mean + d * f;
  r.m2 = a.m2 + b.m2 + d * d * (a.n * f);
  return r;
}

// Block-wide merge of per-thread moments into per-group moments, fixed order (binary tree over the
// nt / q threads that share a channel quad, then the group's quads left to right).
// s_part: [nt] per-thread moments (clobbered); result for group g in s_out[g], g < n_groups.
__device__ __forceinline__ void block_group_moments(Moments *s_part, Moments *s_out, int tid, int nt, int q, int qpg, int n_groups) {
  const int k = tid / q, K = nt / q;
  for (int stride = 1; stride < K; stride <<= 1) {
    __syncthreads();
    if ((k & (2 * stride - 1)) == 0 && k + stride < K) s_part[tid] = merge(s_part[tid], s_part[tid + stride * q]);
  }
  __syncthreads();
  if (tid < n_groups) {
    Moments acc = s_part[tid * qpg];
    for (int t = 1; t < qpg; ++t) acc = merge(acc, s_part[tid * qpg + t]);
    s_out[tid] = acc;
  }
}

// Same result layout, for q | 32 and nt % 32 == 0 (the common shapes): the lanes of a warp that share
// a channel quad merge by shuffles (lower lane = left operand: fixed order, no barrier), one
// __syncthreads publishes the per-warp partials, warp 0 finishes.  The barrier-heavy tree above was
// the top stall of the kernel (ncu: 40 % of warp samples waiting at barriers).
__device__ __forceinline__ void block_group_moments_shfl(Moments m, Moments *s_wpart, Moments *s_out, int tid, int nt, int q, int qpg,
                                                         int n_groups) {
  const int lane = tid & 31, warp = tid >> 5, n_warps = nt >> 5;
  for (int off = q; off < 32; off <<= 1) {
    Moments o;
    o.n = __shfl_xor_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_xor_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, off);
    m = (lane & off) ? merge(o, m) : merge(m, o);
  }
  if (lane < q) s_wpart[warp * q + lane] = m;
  __syncthreads();
  if (warp == 0) {
    if (lane < q) {
      Moments acc = s_wpart[lane];
      for (int w = 1; w < n_warps; ++w) acc = merge(acc, s_wpart[w * q + lane]);
      s_wpart[lane] = acc;
    }
    __syncwarp();
    if (lane < n_groups) {
      Moments acc = s_wpart[lane * qpg];
      for (int t = 1; t < qpg; ++t) acc = merge(acc, s_wpart[lane * qpg + t]);
      s_out[lane] = acc;
    }
  }
}

// SiLU with the SFU exponential and reciprocal: relative error ~1e-6, far below the TF32 convolutions around it
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

constexpr int kGnMaxThreads = 256;

__global__ void __launch_bounds__(kGnMaxThreads) groupnorm_nhwc_kernel(GnArgs a) {
  __shared__ Moments s_part[kGnMaxThreads];
  __shared__ float s_mean[32], s_rstd[32];       // per group of this CTA

  const int b = blockIdx.y;
  const int c0 = blockIdx.x * a.cblk;
  const int q = a.cblk >> 2;                      // float4 quads per pixel in this CTA's channel block
  const int tid = threadIdx.x;                    // blockDim.x is a multiple of q
  const int cq = tid % q;                         // this thread's quad: channels c0 + 4 cq .. + 3
  const int prow = tid / q, pstride = blockDim.x / q;
  const int c = c0 + cq * 4;
  const int g_local = (cq * 4) / a.cpg;
  const int n_groups = a.cblk / a.cpg;

  const size_t base = ((size_t)b * a.HW) * a.C + c;
  const float4 *x4 = reinterpret_cast<const float4 *>(a.x + base);
  const float4 *r4 = a.res ? reinterpret_cast<const float4 *>(a.res + base) : nullptr;
  float4 *s4 = a.sum_out ? reinterpret_cast<float4 *>(a.sum_out + base) : nullptr;
  const int rowq = a.C >> 2;                      // float4 stride between pixels
  float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.add_bc) add = __ldg(reinterpret_cast<const float4 *>(a.add_bc + (size_t)b * a.add_stride + c));

  // ---- pass 1: s = x (+ res) (+ add), shifted sums -------------------------------------------
  // shifted sums over at most 64 pixels at a time, folded into running moments (Chan): a thread walks up to
  // thousands of pixels at 128^2 / 256^2 and one long fp32 sum would lose ~sqrt(n) ulps
  float shift = 0.f, sum = 0.f, sq = 0.f, cnt = 0.f;
  bool first = true;
  Moments run = {0.f, 0.f, 0.f};
  int in_chunk = 0;
  for (int p = prow; p < a.HW; p += pstride * 4) {
    if (in_chunk == 16) {
      Moments c;
      c.n = cnt;
      c.mean = shift + sum / cnt;
      c.m2 = fmaxf(sq - sum * sum / cnt, 0.f);
      run = merge(run, c);
      sum = sq = cnt = 0.f;
      first = true;
      in_chunk = 0;
    }
    ++in_chunk;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = p + u * pstride;
      if (pp < a.HW) {
        v[u] = __ldg(x4 + (size_t)pp * rowq);
        if (r4) {
          const float4 w = __ldg(r4 + (size_t)pp * rowq);
          v[u].x += w.x; v[u].y += w.y; v[u].z += w.z; v[u].w += w.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = p + u * pstride;
      if (pp < a.HW) {
        float4 s = v[u];
        s.x += add.x; s.y += add.y; s.z += add.z; s.w += add.w;
        if (s4) s4[(size_t)pp * rowq] = s;
        if (first) { shift = s.x; first = false; }
        const float d0 = s.x - shift, d1 = s.y - shift, d2 = s.z - shift, d3 = s.w - shift;
        sum += (d0 + d1) + (d2 + d3);
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        cnt += 4.f;
      }
    }
  }
  Moments m;
  m.n = cnt;
  m.mean = cnt > 0.f ? shift + sum / cnt : 0.f;
  m.m2 = cnt > 0.f ? fmaxf(sq - sum * sum / cnt, 0.f) : 0.f;
  m = merge(run, m);
  s_part[tid] = m;
  __shared__ Moments s_grp[32];
  block_group_moments(s_part, s_grp, tid, blockDim.x, q, a.cpg >> 2, n_groups);
  if (tid < n_groups) {
    s_mean[tid] = s_grp[tid].mean;
    s_rstd[tid] = rsqrtf(s_grp[tid].m2 / s_grp[tid].n + a.eps);
  }
  __syncthreads();

  // ---- pass 2: normalise + affine + activation (input re-read: L2-resident) ------------------
  const float mean = s_mean[g_local], rstd = s_rstd[g_local];
  const float4 w = __ldg(reinterpret_cast<const float4 *>(a.weight + c));
  const float4 bi = __ldg(reinterpret_cast<const float4 *>(a.bias + c));
  const float4 sc = make_float4(rstd * w.x, rstd * w.y, rstd * w.z, rstd * w.w);
  float4 *y4 = reinterpret_cast<float4 *>(a.y + base);
  const float4 *in4 = s4 ? reinterpret_cast<const float4 *>(s4) : x4;
  const bool recompute = (s4 == nullptr);
  for (int p = prow; p < a.HW; p += pstride * 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = p + u * pstride;
      if (pp < a.HW) {
        v[u] = in4[(size_t)pp * rowq];
        if (recompute && r4) {
          const float4 t = __ldg(r4 + (size_t)pp * rowq);
          v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = p + u * pstride;
      if (pp < a.HW) {
        float4 s = v[u];
        if (recompute) { s.x += add.x; s.y += add.y; s.z += add.z; s.w += add.w; }
        float4 o;
        o.x = (s.x - mean) * sc.x + bi.x;
        o.y = (s.y - mean) * sc.y + bi.y;
        o.z = (s.z - mean) * sc.z + bi.z;
        o.w = (s.w - mean) * sc.w + bi.w;
        if (a.silu) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
        y4[(size_t)pp * rowq] = o;
      }
    }
  }
}

// ---- cluster variant (the fast path) ----------------------------------------------------------
// A thread-block cluster of P CTAs shares one (sample, channel block): CTA `rank` stages its slab of
// pixels in SHARED MEMORY with cp.async (every byte of the slab is in flight at once -- the
// single-CTA kernel above is latency-bound: it keeps ~16 KiB in flight per SM), computes the
// moments of its slab, the P partial moments are merged in rank order through distributed shared
// memory, and the slab is normalised straight out of shared memory.  HBM traffic: one read, one
// write.  Requires res == NULL and sum_out == NULL (the register kernel handles those).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

// PERSISTENT: a cluster walks the work items (sample, channel block) item = cluster id + k * #clusters.
// While a slab is normalised and stored, every quad a thread has consumed is immediately refilled
// with the NEXT item's data by cp.async (same thread owns the same slab slots in every phase), so
// the HBM reads of item i+1 overlap the writes of item i; without this all resident CTAs moved in
// lock step (load, compute, store) and DRAM sat idle two thirds of the time (ncu: 33 %).
// kCluster = false: P == 1 (small activations), launched without a cluster; block barriers only.
// kPiped (the cluster launches): the slab moves in four cp.async commit groups, see below; the small single-CTA launches
// have one to four iterations per thread and keep one group (the extra waits and commits cost them 10-15 %).
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool kCluster, bool kPiped = kCluster>
__global__ void __launch_bounds__(kGnMaxThreads) groupnorm_nhwc_cluster_kernel(GnArgs a, int P, int ppc, int n_items) {
  constexpr int NG = kPiped ? 4 : 1;                // commit groups per slab
  extern __shared__ __align__(16) float4 tile[];   // [pixels of this CTA][q] channel quads
  __shared__ Moments s_part[kGnMaxThreads];
  __shared__ Moments s_grp[2][32];                  // this CTA's moments per group, double-buffered by item parity
  __shared__ float s_mean[32], s_rstd[32];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = kCluster ? (int)cluster.block_rank() : 0;
  const int cluster_id = blockIdx.x / P, n_clusters = gridDim.x / P;
  const int n_cblk = a.C / a.cblk;
  const int q = a.cblk >> 2;
  const int tid = threadIdx.x, nt = blockDim.x;     // nt is a multiple of q
  const int px0 = rank * ppc;
  const int npx = max(0, min(ppc, a.HW - px0));
  const int n_quads = npx * q;
  const int rowq = a.C >> 2;
  // nt is a multiple of q: thread tid always handles quad cq = tid % q, pixels tid / q + k (nt / q)
  const int cq = tid % q;
  const int prow = tid / q, pstep = nt / q;
  const int g_local = (cq * 4) / a.cpg;
  const int n_groups = a.cblk / a.cpg;

  // source of this thread's quad for work item `item` (x, or x2 for the tail of a concatenated input)
  auto src_of = [&](int item, const float4 *&src, size_t &step) {
    const int b = item / n_cblk, c0 = (item - b * n_cblk) * a.cblk;
    const int cg0 = c0 + cq * 4;
    const bool second = a.x2 != nullptr && cg0 >= a.C1;
    const int srow = second ? a.C - a.C1 : a.C1;
    const float *sbase = second ? a.x2 + (cg0 - a.C1) : a.x + cg0;
    src = reinterpret_cast<const float4 *>(sbase + ((size_t)b * a.HW + px0 + prow) * srow);
    step = (size_t)pstep * (srow >> 2);
  };

  // A thread's slots (i = tid + k nt) are loaded, read and refilled by that thread alone, in k order, so the slab is
  // pipelined in four commit groups: the moments of the first quarter start while the later quarters are still landing
  // (one wait for the whole slab exposed the full HBM latency once per item: 12 % of the warp samples in ncu).
  const int kq = ((n_quads + nt - 1) / nt + NG - 1) / NG > 0 ? ((n_quads + nt - 1) / nt + NG - 1) / NG : 1;      // iterations per group

  int item = cluster_id;
  if (item < n_items) {
    const float4 *src;
    size_t step;
    src_of(item, src, step);
    for (int k = 0, i = tid; k < NG * kq; ++k, i += nt, src += step) {
      if (i < n_quads) cp_async16(&tile[i], src);
      if ((k + 1) % kq == 0) asm volatile("cp.async.commit_group;" ::: "memory");
    }
  } else {
#pragma unroll
    for (int g4 = 0; g4 < NG; ++g4) asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // per-item channel vectors (time-embedding add, affine weight / bias): the NEXT item's are requested before this item's
  // cluster barrier, so their L2 round trip never sits in front of a moments pass (6 % of the warp samples in ncu)
  auto vectors_of = [&](int it, float4 &add_o, float4 &w_o, float4 &bi_o) {
    const int bb = it / n_cblk, cc = (it - bb * n_cblk) * a.cblk + cq * 4;
    add_o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.add_bc) add_o = __ldg(reinterpret_cast<const float4 *>(a.add_bc + (size_t)bb * a.add_stride + cc));
    w_o = __ldg(reinterpret_cast<const float4 *>(a.weight + cc));
    bi_o = __ldg(reinterpret_cast<const float4 *>(a.bias + cc));
  };
  float4 add_n, w_n, bi_n;
  if (item < n_items) vectors_of(item, add_n, w_n, bi_n);

  for (int par = 0; item < n_items; item += n_clusters, par ^= 1) {
    const int b = item / n_cblk, c0 = (item - b * n_cblk) * a.cblk;
    const int c = c0 + cq * 4;
    const float4 add = add_n, w = w_n, bi = bi_n;
    // ---- moments of this CTA's slab (s = x + add is written back so the second pass reads s) -----
    // (no block barrier: every slot is touched by one thread only; s_grp / s_mean are ordered by the barriers below)
    float shift = 0.f, sum = 0.f, sq = 0.f, cnt = 0.f;
    cp_async_wait<NG - 1>();
    if (tid < n_quads) shift = tile[tid].x + add.x;
#pragma unroll
    for (int g4 = 0; g4 < NG; ++g4) {
      if (g4 == 1) cp_async_wait<2>();
      if (g4 == 2) cp_async_wait<1>();
      if (g4 == 3) cp_async_wait<0>();
#pragma unroll 4
      for (int k = 0; k < kq; ++k) {
        const int i = tid + (g4 * kq + k) * nt;
        if (i >= n_quads) break;
        float4 s = tile[i];
        s.x += add.x; s.y += add.y; s.z += add.z; s.w += add.w;
        if (a.add_bc) tile[i] = s;
        const float d0 = s.x - shift, d1 = s.y - shift, d2 = s.z - shift, d3 = s.w - shift;
        sum += (d0 + d1) + (d2 + d3);
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        cnt += 4.f;
      }
    }
    Moments m;
    m.n = cnt;
    m.mean = cnt > 0.f ? shift + sum / cnt : 0.f;
    m.m2 = cnt > 0.f ? fmaxf(sq - sum * sum / cnt, 0.f) : 0.f;
    if ((32 % q) == 0 && (nt & 31) == 0) {
      block_group_moments_shfl(m, s_part, s_grp[par], tid, nt, q, a.cpg >> 2, n_groups);
    } else {
      s_part[tid] = m;
      block_group_moments(s_part, s_grp[par], tid, nt, q, a.cpg >> 2, n_groups);
    }
    if (item + n_clusters < n_items) vectors_of(item + n_clusters, add_n, w_n, bi_n);
    // One cluster barrier per item: s_grp is double-buffered, and a CTA can be at most one item ahead of
    // a peer (it blocks at the next barrier), so nobody overwrites moments a peer still has to read.
    if (kCluster) cluster.sync(); else __syncthreads();
    if (tid < n_groups) {
      Moments acc = {0.f, 0.f, 0.f};
      if (kCluster) {
        for (int r = 0; r < P; ++r) {               // rank order: every CTA computes the same bits
          const Moments *remote = cluster.map_shared_rank(&s_grp[par][0], r);
          acc = merge(acc, remote[tid]);
        }
      } else {
        acc = s_grp[par][tid];
      }
      s_mean[tid] = acc.mean;
      s_rstd[tid] = rsqrtf(acc.m2 / acc.n + a.eps);
    }
    __syncthreads();

    // ---- normalise + affine + activation out of shared memory; refill consumed slots with the next item ----
    // y = s * sc + sh with sc = rstd * weight, sh = bias - mean * sc (the form PyTorch's GroupNorm uses too)
    const float mean = s_mean[g_local], rstd = s_rstd[g_local];
    const float4 sc = make_float4(rstd * w.x, rstd * w.y, rstd * w.z, rstd * w.w);
    const float4 sh = make_float4(bi.x - mean * sc.x, bi.y - mean * sc.y, bi.z - mean * sc.z, bi.w - mean * sc.w);
    float4 *y4 = reinterpret_cast<float4 *>(a.y + ((size_t)b * a.HW + px0) * a.C + c0) + (size_t)prow * rowq + cq;
    const size_t ystep = (size_t)pstep * rowq;
    const int next = item + n_clusters;
    const float4 *nsrc = nullptr;
    size_t nstep = 0;
    if (next < n_items) src_of(next, nsrc, nstep);
#pragma unroll
    for (int g4 = 0; g4 < NG; ++g4) {
#pragma unroll 4
      for (int k = 0; k < kq; ++k) {
        const int i = tid + (g4 * kq + k) * nt;
        if (i >= n_quads) break;
        const float4 s = tile[i];
        if (nsrc) {
          cp_async16(&tile[i], nsrc);
          nsrc += nstep;
        }
        float4 o;
        o.x = fmaf(s.x, sc.x, sh.x);
        o.y = fmaf(s.y, sc.y, sh.y);
        o.z = fmaf(s.z, sc.z, sh.z);
        o.w = fmaf(s.w, sc.w, sh.w);
        if (a.silu) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
        *y4 = o;
        y4 += ystep;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  if (kCluster) cluster.sync();                     // nobody leaves while a peer may still read its moments
}

// ---- warp variant (the 2x2 / 4x4 / 8x8 levels) ------------------------------------------------
// At the UNet's low resolutions a work item (sample, 32-channel block) is 0.5-8 KiB: staging it in shared memory and
// meeting at block barriers costs more than moving it (4-10 us per launch for 2-17 MB).  Here ONE WARP owns an item in
// registers: every lane loads its <= KMAX channel quads with independent 16-byte loads, the moments are merged by
// shuffles only (lanes that share a quad, then the quads of a group; lower lane = left operand: fixed order), and the
// lanes normalise and store what they hold.  No shared memory, no barrier, all items of a launch in flight at once.
template <int KMAX>
__global__ void __launch_bounds__(256) groupnorm_nhwc_warp_kernel(GnArgs a, int n_items) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const int n_cblk = a.C >> 5;
  const int b = item / n_cblk, c0 = (item - b * n_cblk) << 5;
  const int cq = lane & 7, prow = lane >> 3;      // 8 quads per pixel, 4 pixels per pass
  const int c = c0 + cq * 4;
  const bool second = a.x2 != nullptr && c >= a.C1;
  const int srow = second ? a.C - a.C1 : a.C1;
  const float *sbase = second ? a.x2 + (c - a.C1) : a.x + c;
  const float4 *src = reinterpret_cast<const float4 *>(sbase + ((size_t)b * a.HW + prow) * srow);
  const size_t sstep = (size_t)srow;             // float4 units between passes: 4 pixels x srow floats / 4
  float4 v[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (prow + 4 * k < a.HW) v[k] = __ldg(src + (size_t)k * sstep);
  float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.add_bc) add = __ldg(reinterpret_cast<const float4 *>(a.add_bc + (size_t)b * a.add_stride + c));
  const float4 w = __ldg(reinterpret_cast<const float4 *>(a.weight + c));
  const float4 bi = __ldg(reinterpret_cast<const float4 *>(a.bias + c));

  float shift = 0.f, sum = 0.f, sq = 0.f, cnt = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (prow + 4 * k < a.HW) {
      v[k].x += add.x; v[k].y += add.y; v[k].z += add.z; v[k].w += add.w;
      if (k == 0) shift = v[0].x;
      const float d0 = v[k].x - shift, d1 = v[k].y - shift, d2 = v[k].z - shift, d3 = v[k].w - shift;
      sum += (d0 + d1) + (d2 + d3);
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      cnt += 4.f;
    }
  Moments m;
  m.n = cnt;
  m.mean = cnt > 0.f ? shift + sum / cnt : 0.f;
  m.m2 = cnt > 0.f ? fmaxf(sq - sum * sum / cnt, 0.f) : 0.f;
  const int qpg = a.cpg >> 2;                    // quads per group: 1, 2, 4 or 8
  // lanes that share a quad (xor 8, 16), then the quads of a group (xor 1 .. qpg / 2)
  for (int off = 8; off < 32; off <<= 1) {
    Moments o;
    o.n = __shfl_xor_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_xor_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, off);
    m = (lane & off) ? merge(o, m) : merge(m, o);
  }
  for (int off = 1; off < qpg; off <<= 1) {
    Moments o;
    o.n = __shfl_xor_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_xor_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, off);
    m = (lane & off) ? merge(o, m) : merge(m, o);
  }
  const float rstd = rsqrtf(m.m2 / m.n + a.eps);
  const float4 sc = make_float4(rstd * w.x, rstd * w.y, rstd * w.z, rstd * w.w);
  const float4 sh = make_float4(bi.x - m.mean * sc.x, bi.y - m.mean * sc.y, bi.z - m.mean * sc.z, bi.w - m.mean * sc.w);
  float4 *y4 = reinterpret_cast<float4 *>(a.y + ((size_t)b * a.HW + prow) * a.C + c);
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (prow + 4 * k < a.HW) {
      float4 o;
      o.x = fmaf(v[k].x, sc.x, sh.x);
      o.y = fmaf(v[k].y, sc.y, sh.y);
      o.z = fmaf(v[k].z, sc.z, sh.z);
      o.w = fmaf(v[k].w, sc.w, sh.w);
      if (a.silu) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
      y4[(size_t)k * a.C] = o;                    // 4 pixels further: 4 x C floats = C float4
    }
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// K6: out = (a [+ bias_a[c]]) + (b + bias_b[c]) on NHWC activations: the biases of conv_shortcut /
// conv2 and the residual add of a ResnetBlock2D in one pass (PyTorch: a bias-add pass inside each
// convolution, then the add); same association, so bit-identical to that sequence.
__global__ void __launch_bounds__(256) add_bias_nhwc_kernel(const float4 *__restrict__ a, const float4 *__restrict__ a2,
                                                            const float4 *__restrict__ bias_a, const float4 *__restrict__ b,
                                                            const float4 *__restrict__ bias_b, float4 *__restrict__ out, size_t n4,
                                                            int cq) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = a[i];
    if (a2) {
      const float4 u = __ldg(a2 + i);
      x.x = __fadd_rn(x.x, u.x); x.y = __fadd_rn(x.y, u.y); x.z = __fadd_rn(x.z, u.z); x.w = __fadd_rn(x.w, u.w);
    }
    const float4 y = __ldg(b + i), z = __ldg(bias_b + (i % cq));
    if (bias_a) {
      const float4 w = __ldg(bias_a + (i % cq));
      x.x = __fadd_rn(x.x, w.x); x.y = __fadd_rn(x.y, w.y); x.z = __fadd_rn(x.z, w.z); x.w = __fadd_rn(x.w, w.w);
    }
    float4 o;
    o.x = __fadd_rn(x.x, __fadd_rn(y.x, z.x));
    o.y = __fadd_rn(x.y, __fadd_rn(y.y, z.y));
    o.z = __fadd_rn(x.z, __fadd_rn(y.z, z.z));
    o.w = __fadd_rn(x.w, __fadd_rn(y.w, z.w));
    out[i] = o;
  }
}

cudaError_t launch_add_bias_nhwc(const float *a, const float *a2, const float *bias_a, const float *b, const float *bias_b,
                                 float *out, size_t n, int C, cudaStream_t s) {
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_bias_nhwc_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(a2),
                                                        reinterpret_cast<const float4 *>(bias_a),
                                                        reinterpret_cast<const float4 *>(b), reinterpret_cast<const float4 *>(bias_b),
                                                        reinterpret_cast<float4 *>(out), n4, C / 4);
  return cudaGetLastError();
}

cudaError_t launch_groupnorm_nhwc(const float *x, const float *x2, int C1, const float *res, const float *add_bc, int add_stride,
                                  const float *weight, const float *bias, float *sum_out, float *y, int B, int C, int HW,
                                  int groups, float eps, int silu, cudaStream_t s) {
  GnArgs a;
  a.add_stride = add_stride;
  a.x2 = x2;
  a.C1 = x2 ? C1 : C;
  if (x2 && (res || sum_out)) return cudaErrorInvalidValue;      // two sources: cluster kernel only
  a.x = x; a.res = res; a.add_bc = add_bc; a.weight = weight; a.bias = bias; a.sum_out = sum_out; a.y = y;
  a.B = B; a.C = C; a.HW = HW; a.cpg = C / groups; a.eps = eps; a.silu = silu;
  int cblk = a.cpg / gcd_i(a.cpg, 32) * 32;        // lcm(cpg, 32): whole groups and whole 128-byte lines
  if (cblk > C || C % cblk != 0) cblk = C;         // odd shapes: one CTA per sample takes all channels
  a.cblk = cblk;
  const int q = cblk / 4;
  if (q > kGnMaxThreads || cblk / a.cpg > 32) return cudaErrorInvalidValue;
  const int threads = kGnMaxThreads / q * q;
  const int cblk0 = cblk;                          // the register kernel below keeps whole 128-byte lines

  if (!res && !sum_out && cblk == 32 && HW <= 64) {
    // warp path: the 2x2 / 4x4 / 8x8 levels, one warp per (sample, 32-channel block)
    static thread_local int warp_path = -1;    // BNDM_GN_WARP=0 switches it off (A/B measurements)
    if (warp_path < 0) { const char *e = getenv("BNDM_GN_WARP"); warp_path = e ? atoi(e) : 1; }
    const int qpg = a.cpg >> 2;
    if (warp_path && (qpg == 1 || qpg == 2 || qpg == 4 || qpg == 8)) {
      const int n_items = B * (C / 32);
      const int blocks = (n_items + 7) / 8;
      if (HW <= 4) groupnorm_nhwc_warp_kernel<1><<<blocks, 256, 0, s>>>(a, n_items);
      else if (HW <= 16) groupnorm_nhwc_warp_kernel<4><<<blocks, 256, 0, s>>>(a, n_items);
      else groupnorm_nhwc_warp_kernel<16><<<blocks, 256, 0, s>>>(a, n_items);
      return cudaGetLastError();
    }
  }
  if (!res && !sum_out) {
    // cluster path: P CTAs per (sample, channel block), slab of HW / P pixels in shared memory
    static thread_local int slab_kb = 0;      // BNDM_GN_SLAB_KB: target slab size per CTA (experiments); default 64
    if (!slab_kb) {
      const char *e = getenv("BNDM_GN_SLAB_KB");
      slab_kb = e ? atoi(e) : 64;
      if (slab_kb < 8 || slab_kb > 192) slab_kb = 64;
    }
    // Channel block of the cluster path.  Whole 128-byte lines (32 channels) make a 64^2 item 512 KiB = 8 CTAs x 64 KiB,
    // and clusters of 8 fit on only 128 of the 148 SMs; at 128^2 the item is 2 MiB = 16 CTAs x 128 KiB, non-portable
    // clusters on one CTA per SM (measured 2.2 TB/s).  Narrower blocks -- 64- or 32-byte row pieces, still whole groups --
    // keep three 64 KiB CTAs per SM and let the clusters shrink: measured (B200, tools/k5_micro.py)
    //   64^2 x 128 ch: 32 ch / P=8 3.96 TB/s, 16 ch / P=4 4.44, 8 ch / P=2 3.94;   128^2 x 128 ch: 32 ch 2.2, 16 ch 3.2, 8 ch / P=8 3.4.
    // Rule: halve towards 4 x 64 KiB while the pieces stay >= 64 bytes, then towards 8 x 64 KiB down to 32-byte pieces.
    static thread_local int force_cblk = -1;   // BNDM_GN_CBLK: force a channel block (experiments)
    if (force_cblk < 0) { const char *e = getenv("BNDM_GN_CBLK"); force_cblk = e ? atoi(e) : 0; }
    // (the same block with or without a second source -- a thread picks its source by its own channel quad -- so
    // the two-source call stays bit-identical to the call on the concatenated tensor)
    auto can_halve = [&](int cb) { return cb % 2 == 0 && (cb / 2) % a.cpg == 0 && (cb / 2) % 8 == 0; };
    if (force_cblk > 0 && force_cblk % a.cpg == 0 && C % force_cblk == 0 && force_cblk % 4 == 0) {
      cblk = force_cblk;
    } else {
      while ((size_t)HW * cblk * 4 > (size_t)4 * slab_kb * 1024 && cblk / 2 >= 16 && can_halve(cblk)) cblk /= 2;
      while ((size_t)HW * cblk * 4 > (size_t)8 * slab_kb * 1024 && cblk / 2 >= 8 && can_halve(cblk)) cblk /= 2;
    }
    a.cblk = cblk;
    const int q = cblk / 4;
    const size_t row_bytes = (size_t)cblk * 4;
    const int p_max = slab_kb < 64 ? 16 : 8;
    int P = 1;
    while (P < p_max && ((size_t)((HW + P - 1) / P) * row_bytes > (size_t)slab_kb * 1024)) P *= 2;
    if ((size_t)((HW + P - 1) / P) * row_bytes > 200 * 1024) P = 16;     // 128^2 images: non-portable cluster size
    const int ppc = (HW + P - 1) / P;
    const size_t smem = (size_t)ppc * row_bytes;
    if (smem <= 200 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(groupnorm_nhwc_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(groupnorm_nhwc_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess && P > 8) e = cudaFuncSetAttribute(groupnorm_nhwc_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      // small slabs: no more thread rows than pixels (shallower merge tree, cheaper CTAs)
      int rows = kGnMaxThreads / q;
      if (rows > ppc) rows = ppc;
      while (rows * q < 32) ++rows;
      if (32 % q == 0) rows = (rows * q + 31) / 32 * 32 / q;      // whole warps: the shuffle merge needs them
      const int nt = rows * q;
      const int n_items = B * (C / cblk);
      if (P == 1) {
        // small activations: no cluster; persistent over min(items, resident CTAs)
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, groupnorm_nhwc_cluster_kernel<false>, nt, smem) != cudaSuccess || per_sm < 1) {
          cudaGetLastError();
          per_sm = 1;
        }
        const int resident = per_sm * 148;
        groupnorm_nhwc_cluster_kernel<false><<<n_items < resident ? n_items : resident, nt, smem, s>>>(a, 1, ppc, n_items);
        return cudaGetLastError();
      }
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3((unsigned)nt);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)P;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      // persistent grid: as many clusters as can be resident at once (cached per launch shape)
      static thread_local int cache_key[4] = {0, 0, 0, -1}, cache_val = 0;     // per host thread (and device): no shared mutable state
      int max_clusters = 0;
      int dev_id = 0;
      cudaGetDevice(&dev_id);
      if (cache_key[0] == P && cache_key[1] == nt && cache_key[2] == (int)smem && cache_key[3] == dev_id) {
        max_clusters = cache_val;
      } else {
        cfg.gridDim = dim3((unsigned)P, 1, 1);
        if (cudaOccupancyMaxActiveClusters(&max_clusters, groupnorm_nhwc_cluster_kernel<true>, &cfg) != cudaSuccess || max_clusters < 1) {
          cudaGetLastError();
          max_clusters = 148 / P > 0 ? 148 / P : 1;
        }
        cache_key[0] = P; cache_key[1] = nt; cache_key[2] = (int)smem; cache_key[3] = dev_id; cache_val = max_clusters;
      }
      int n_clusters = n_items < max_clusters ? n_items : max_clusters;
      {
        // Even rounds: a persistent cluster walks ceil(n_items / n_clusters) items, so 256 items on 48 clusters cost 6
        // rounds with a sixth of the machine idle in the last one.  Take the FEWEST clusters that still finish in the
        // same number of rounds (fewer clusters = fewer cluster barriers in flight, same span) -- or, when
        // BNDM_GN_EVEN=2, the largest count that divides the items evenly if that loses no more than a quarter of the
        // resident clusters.
        static thread_local int even_mode = -1;
        if (even_mode < 0) { const char *e = getenv("BNDM_GN_EVEN"); even_mode = e ? atoi(e) : 0; }
        const int rounds = (n_items + n_clusters - 1) / n_clusters;
        if (even_mode == 1) n_clusters = (n_items + rounds - 1) / rounds;
        if (even_mode == 2)
          for (int c = n_clusters; c >= (3 * n_clusters) / 4 && c >= 1; --c)
            if (n_items % c == 0) { n_clusters = c; break; }
      }
      cfg.gridDim = dim3((unsigned)(P * n_clusters), 1, 1);
      return cudaLaunchKernelEx(&cfg, groupnorm_nhwc_cluster_kernel<true>, a, P, ppc, n_items);
    }
  }
  if (x2) return cudaErrorInvalidValue;                           // slab does not fit: caller concatenates first
  a.cblk = cblk0;
  dim3 grid(C / cblk0, B);
  groupnorm_nhwc_kernel<<<grid, threads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace bndm

// K7: self-attention of the UNet's attention blocks at their tiny sizes (diffusers Attention with
// attention_head_dim = 8 at 4x4 / 2x2 resolution: T = 16 or 4 tokens, C / 8 heads; iadb_bn.py:216,222).
// qkv: [B][T][3C] (q | k | v along the last axis, head h = channels 8h .. 8h+7), out: [B][T][C].
// One thread per (sample, head, query token): scores against the T keys in registers, softmax,
// weighted sum of the values.  PyTorch dispatches these shapes to a generic flash-attention kernel
// that takes ~80 us per call at B=64; this is a few microseconds of L2-resident traffic.
namespace bndm {

constexpr int kAttnMaxT = 64;

template <int D>
__global__ void __launch_bounds__(256) attention_small_kernel(const float *__restrict__ qkv, float *__restrict__ out, int B, int T,
                                                              int C, float scale) {
  const int heads = C / D;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // ((b * T + t) * heads + h)
  if (idx >= (int64_t)B * T * heads) return;
  const int h = (int)(idx % heads);
  const int64_t bt = idx / heads;
  const int b = (int)(bt / T);
  const float *q = qkv + bt * 3 * C + h * D;
  const float *kbase = qkv + (int64_t)b * T * 3 * C + C + h * D;
  float qv[D];
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(q + d));
    qv[d] = v.x; qv[d + 1] = v.y; qv[d + 2] = v.z; qv[d + 3] = v.w;
  }
  float s[kAttnMaxT];
  float m = -INFINITY;
  for (int j = 0; j < T; ++j) {
    const float *k = kbase + (int64_t)j * 3 * C;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(k + d));
      acc += qv[d] * v.x + qv[d + 1] * v.y + qv[d + 2] * v.z + qv[d + 3] * v.w;
    }
    s[j] = acc * scale;
    m = fmaxf(m, s[j]);
  }
  float denom = 0.f;
  for (int j = 0; j < T; ++j) {
    s[j] = expf(s[j] - m);
    denom += s[j];
  }
  const float inv = 1.0f / denom;
  float o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  for (int j = 0; j < T; ++j) {
    const float *v = kbase + C + (int64_t)j * 3 * C;
    const float p = s[j] * inv;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      const float4 w = __ldg(reinterpret_cast<const float4 *>(v + d));
      o[d] += p * w.x; o[d + 1] += p * w.y; o[d + 2] += p * w.z; o[d + 3] += p * w.w;
    }
  }
  float *dst = out + bt * C + h * D;
#pragma unroll
  for (int d = 0; d < D; d += 4) *reinterpret_cast<float4 *>(dst + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
}

cudaError_t launch_attention_small(const float *qkv, float *out, int B, int T, int C, int head_dim, cudaStream_t s) {
  if (head_dim != 8 || T > kAttnMaxT || C % head_dim != 0) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)B * T * (C / head_dim);
  attention_small_kernel<8><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(qkv, out, B, T, C, 1.0f / sqrtf((float)head_dim));
  return cudaGetLastError();
}

}  // namespace bndm

// K8: nearest-neighbour 2x upsampling of a channels-last activation (diffusers Upsample2D in front of its
// 3x3 convolution): y[b][2i+a][2j+c][:] = x[b][i][j][:].  One thread per input channel quad: one 16-byte
// load, four 16-byte stores, all coalesced (PyTorch's upsample_nearest2d_nhwc kernel moves the same bytes
// at under 1 TB/s).
namespace bndm {

__global__ void __launch_bounds__(256) upsample2x_nhwc_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, int H, int W,
                                                              int cq, size_t n_in) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const size_t pix = i / cq;                    // ((b * H + h) * W + w)
    const int c = (int)(i - pix * cq);
    const int w = (int)(pix % W);
    const size_t bh = pix / W;                    // b * H + h
    const float4 v = __ldg(x + i);
    float4 *o = y + ((bh * 2) * (size_t)(2 * W) + 2 * w) * cq + c;      // output row 2h of image b (2H rows per image)
    o[0] = v;
    o[cq] = v;
    o += (size_t)(2 * W) * cq;
    o[0] = v;
    o[cq] = v;
  }
}

cudaError_t launch_upsample2x_nhwc(const float *x, float *y, int B, int H, int W, int C, cudaStream_t s) {
  const size_t n_in = (size_t)B * H * W * (C / 4);
  size_t blocks = (n_in + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_nhwc_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(y), H, W,
                                                          C / 4, n_in);
  return cudaGetLastError();
}

}  // namespace bndm
