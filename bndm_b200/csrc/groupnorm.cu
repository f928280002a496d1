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
#include "common.cuh"

namespace bndm {

struct GnArgs {
  const float *x, *res, *add_bc, *weight, *bias;
  float *sum_out, *y;
  int B, C, HW, cpg;       // cpg = channels per group (multiple of 4)
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
  r.mean = a.mean + d * (b.n / n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
  return r;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

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
  if (a.add_bc) add = __ldg(reinterpret_cast<const float4 *>(a.add_bc + (size_t)b * a.C + c));

  // ---- pass 1: s = x (+ res) (+ add), shifted sums -------------------------------------------
  float shift = 0.f, sum = 0.f, sq = 0.f, cnt = 0.f;
  bool first = true;
  for (int p = prow; p < a.HW; p += pstride * 4) {
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
  s_part[tid] = m;
  __syncthreads();
  // fixed-order merge in two levels: quad column (threads with the same cq), then the group's quads
  __shared__ Moments s_col[kGnMaxThreads];
  if (tid < q) {
    Moments acc = s_part[tid];
    for (int t = tid + q; t < (int)blockDim.x; t += q) acc = merge(acc, s_part[t]);
    s_col[tid] = acc;
  }
  __syncthreads();
  if (tid < n_groups) {
    const int qpg = a.cpg >> 2;                   // quads per group
    Moments acc = s_col[tid * qpg];
    for (int t = 1; t < qpg; ++t) acc = merge(acc, s_col[tid * qpg + t]);
    s_mean[tid] = acc.mean;
    s_rstd[tid] = rsqrtf(acc.m2 / acc.n + a.eps);
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

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

cudaError_t launch_groupnorm_nhwc(const float *x, const float *res, const float *add_bc, const float *weight,
                                  const float *bias, float *sum_out, float *y, int B, int C, int HW, int groups, float eps,
                                  int silu, cudaStream_t s) {
  GnArgs a;
  a.x = x; a.res = res; a.add_bc = add_bc; a.weight = weight; a.bias = bias; a.sum_out = sum_out; a.y = y;
  a.B = B; a.C = C; a.HW = HW; a.cpg = C / groups; a.eps = eps; a.silu = silu;
  int cblk = a.cpg / gcd_i(a.cpg, 32) * 32;        // lcm(cpg, 32): whole groups and whole 128-byte lines
  if (cblk > C || C % cblk != 0) cblk = C;         // odd shapes: one CTA per sample takes all channels
  a.cblk = cblk;
  const int q = cblk / 4;
  if (q > kGnMaxThreads || cblk / a.cpg > 32) return cudaErrorInvalidValue;
  const int threads = kGnMaxThreads / q * q;
  dim3 grid(C / cblk, B);
  groupnorm_nhwc_kernel<<<grid, threads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace bndm
