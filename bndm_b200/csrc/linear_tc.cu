// K9: the fp32 linears of the UNet's attention blocks (diffusers Attention to_q / to_k / to_v / to_out at the 4x4 level
// of the model the reference samples with, iadb_bn.py:205-282 / :319) as a 3xTF32 tcgen05 GEMM
//
//     out[m][n] = sum_k a[m][k] * w[n][k] (+ bias[n])          a: [M][K] fp32,  w: [N][K] fp32 (torch.nn.Linear layout)
//
// torch keeps matmuls in fp32 unless told otherwise (torch.backends.cuda.matmul.allow_tf32 = False, as the reference runs),
// so these go to cuBLAS' SIMT sgemm: 40 us for the fused q/k/v projection at batch 64 (M = 1024, N = 1536, K = 512),
// 5 % of a forward.  Here both operands are split  v = hi + lo  with hi = the 19 bits the tensor core reads from an fp32
// word (it ignores the 13 low mantissa bits: the raw tile IS the hi operand) and lo = v - hi (exact in fp32), and
//     a w ~= a_hi w_hi + (a_hi w_lo + a_lo w_hi)            (the dropped lo x lo term is ~2^-22 relative)
// runs as three TF32 MMAs per 8 k: fp32-grade results (measured against fp64 in the tests) at tensor-core speed.
//
// One CTA per 128 (output features, TMEM lanes) x 128 (rows m, TMEM columns) tile; the raw w and a tiles of 128 x 32 k
// arrive by TMA (2-D tensor maps, 128-byte swizzle, rows past the end zero-filled) through a 3-stage ring, four converter
// warps write the lo tiles next to them in shared memory (the kernel is bound by the bytes TMA delivers to an SM, so the lo
// parts are never read from memory), the large main sums and the small corrections accumulate in separate TMEM columns and
// meet in fp32 registers in the epilogue.  Measured (tools/linear_probe.py, B200): 17.5 us for every shape of the UNet
// (M = 256..1024, N = 512..1536, K = 512: 44.7 / 19.8 / 19.0 us in torch).  The bound is SHARED-MEMORY BANDWIDTH, as for K1b:
// per 32-k stage the three MMAs of each of the 4 k-steps read 8 KiB of operands each (96 KiB), the converter reads and
// writes 64 KiB, TMA writes 32 KiB -- 192 KiB at 128 B/clk = 1536 cycles against 768 cycles of tensor work; a 4-deep raw ring
// with a separate lo ring (more loads in flight) and a per-CTA k offset (L2 hot-spotting) changed nothing, which is what
// a shared-memory bound predicts.  max|err| against fp64 ~2x torch's fp32 GEMM (64 k-steps accumulate in TMEM).
// CTA: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue (one TMEM lane quarter each),
// warps 6..9 = converter.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "umma.cuh"

namespace bndm {

constexpr int kLtThreads = 320;
constexpr int kLtStages = 3;
constexpr uint32_t kLtTile = kBlk * kStageK * 4;          // 16 KiB: 128 rows x 32 k
constexpr uint32_t kLtStageBytes = 4 * kLtTile;           // w | a | w_lo | a_lo  (the first two by TMA, the last two by the converter)
constexpr uint32_t kLtTmemCols = 256;                     // main | corrections

struct LtArgs {
  float *out;
  const float *bias;
  int M, N, K;
  int n_nt;                 // output-feature tiles (blockIdx.x % n_nt)
};

__global__ void __launch_bounds__(kLtThreads, 1)
linear_tc_kernel(const LtArgs a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a) {
  extern __shared__ __align__(1024) uint8_t lt_smem[];
  uint8_t *base = lt_smem + ((1024u - (smem_u32(lt_smem) & 1023u)) & 1023u);      // keeps the shared address space
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + (size_t)kLtStages * kLtStageBytes);
  uint64_t *empty_bar = full_bar + kLtStages;
  uint64_t *conv_bar = empty_bar + kLtStages;     // lo tiles written: the stage is ready for the issuer
  uint64_t *acc_full = conv_bar + kLtStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % a.n_nt, mt = blockIdx.x / a.n_nt;
  const int n_stages = a.K / kStageK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    for (int s = 0; s < kLtStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], 4);                 // one arrival per converter warp
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kLtTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer: the two raw tiles of a stage =================
    for (int it = 0; it < n_stages; ++it) {
      const int st = it % kLtStages;
      const uint32_t ph = (uint32_t)(it / kLtStages) & 1u;
      mbar_wait(&empty_bar[st], ph ^ 1u);
      const uint32_t sa = smem_u32(base + (size_t)st * kLtStageBytes);
      if (elect_one()) {
        const int ks = it;
        mbar_expect_tx(&full_bar[st], 2 * kLtTile);
        tma_load_2d(sa, &map_w, &full_bar[st], ks * kStageK, nt * kBlk, kEvictLast);
        tma_load_2d(sa + kLtTile, &map_a, &full_bar[st], ks * kStageK, mt * kBlk, kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer: D[128 features][128 rows]: main = w a^T, corr = w a_lo^T + w_lo a^T =================
    constexpr uint32_t idesc = umma_idesc(kBlk);
    for (int it = 0; it < n_stages; ++it) {
      const int st = it % kLtStages;
      const uint32_t ph = (uint32_t)(it / kLtStages) & 1u;
      mbar_wait(&conv_bar[st], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(base + (size_t)st * kLtStageBytes);
      if (elect_one()) {
        const uint64_t dW = umma_desc(sa), dA = umma_desc(sa + kLtTile), dWl = umma_desc(sa + 2 * kLtTile), dAl = umma_desc(sa + 3 * kLtTile);
#pragma unroll
        for (int kk = 0; kk < kStageK / kUmmaK; ++kk) {
          const uint64_t adv = (uint64_t)((kk * kUmmaK * 4) >> 4);        // +32 B inside the swizzle row
          const uint32_t acc = (it | kk) != 0;
          umma_tf32(tmem_base, dW + adv, dA + adv, idesc, acc);            // hi x hi        -> main
          umma_tf32(tmem_base + kBlk, dW + adv, dAl + adv, idesc, acc);    // hi x lo        -> corrections
          umma_tf32(tmem_base + kBlk, dWl + adv, dA + adv, idesc, 1u);     // lo x hi        -> corrections
        }
        umma_commit(&empty_bar[st]);        // frees the stage when these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 6) {
    // ================= converter: lo = v - trunc_tf32(v) for both raw tiles, 16 x 16 bytes per thread per stage =================
    // (element-wise on the swizzled image: the lo tile has the same layout as its raw tile)
    const int t = threadIdx.x - 192;
    for (int it = 0; it < n_stages; ++it) {
      const int st = it % kLtStages;
      const uint32_t ph = (uint32_t)(it / kLtStages) & 1u;
      mbar_wait(&full_bar[st], ph);
      const float4 *src = reinterpret_cast<const float4 *>(base + (size_t)st * kLtStageBytes) + t;
      float4 *dst = reinterpret_cast<float4 *>(base + (size_t)st * kLtStageBytes + 2 * kLtTile) + t;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = src[(half * 8 + i) * 128];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 lo;
          lo.x = __fsub_rn(v[i].x, __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u));
          lo.y = __fsub_rn(v[i].y, __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u));
          lo.z = __fsub_rn(v[i].z, __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u));
          lo.w = __fsub_rn(v[i].w, __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u));
          dst[(half * 8 + i) * 128] = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_bar[st]);
    }
  } else {
    // ================= epilogue: out[m][n] = main + corr (+ bias[n]); 32 lanes = 32 consecutive n = 128 bytes =================
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int n = nt * kBlk + q * 32 + lane;
    const float bias = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.0f;
    mbar_wait(acc_full, 0u);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int m0 = mt * kBlk;
    for (int j = 0; j < kBlk; j += 16) {
      uint32_t v[16], c[16];
      tmem_ld16(taddr + (uint32_t)j, v);
      tmem_ld16(taddr + (uint32_t)(kBlk + j), c);
      tmem_ld_wait();
      if (n < a.N) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int m = m0 + j + e;
          if (m < a.M) a.out[(size_t)m * a.N + n] = __fadd_rn(__fadd_rn(__uint_as_float(v[e]), __uint_as_float(c[e])), bias);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kLtTmemCols);
  }
}

// [rows][K] fp32 row-major, box = 128 rows x 32 k, 128-byte swizzle, rows past the end read as zero
static bool lt_map(CUtensorMap *m, const float *ptr, int rows, int K) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)kStageK, (cuuint32_t)kBlk};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// a: [M][K], w: [N][K], bias: [N] or null, out: [M][N]
cudaError_t launch_linear_tc(const float *a, const float *w, const float *bias, float *out, int M, int N, int K, cudaStream_t s) {
  if (M < 1 || N < 1 || K < kStageK || K % kStageK != 0 || N % 4 != 0) return cudaErrorNotSupported;
  CUtensorMap map_w, map_a;
  if (!lt_map(&map_w, w, N, K) || !lt_map(&map_a, a, M, K)) {
    set_error("linear_tc: cuTensorMapEncodeTiled failed (M %d, N %d, K %d)", M, N, K);
    return cudaErrorInvalidValue;
  }
  LtArgs g;
  g.out = out; g.bias = bias; g.M = M; g.N = N; g.K = K;
  g.n_nt = (N + kBlk - 1) / kBlk;
  const int n_mt = (M + kBlk - 1) / kBlk;
  const size_t smem = (size_t)kLtStages * kLtStageBytes + 1024 + (3 * kLtStages + 1) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  linear_tc_kernel<<<(unsigned)(g.n_nt * n_mt), kLtThreads, smem, s>>>(g, map_w, map_a);
  return cudaGetLastError();
}

}  // namespace bndm
