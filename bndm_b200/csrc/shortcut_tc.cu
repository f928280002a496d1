// K10: the tail of a diffusers ResnetBlock2D whose input and output channel counts differ (every resnet of the up blocks,
// whose input is the concatenation [h | skip]; the model the reference samples with, iadb_bn.py:205-282 / :319):
//
//     out[m][n] = sum_{k < C1} w[n][k] x[m][k] + sum_{k < C2} w[n][C1 + k] x2[m][k]  +  h2[m][n]  +  bias[n]
//                 \------------- conv_shortcut, a 1x1 convolution of cat(x, x2) ------/    conv2     both biases
//
// on channels-last activations (m = (b, h, w) rows).  As separate steps this is two cuDNN GEMMs (x and x2 are never
// concatenated) that each write a full activation, plus a four-input add pass (K6): 8 passes over a 134 MB tensor at 64^2 x 128
// channels, ~200 us per block.  Here the 1x1 convolution runs on tcgen05 (TF32 inputs, fp32 accumulation, like the cuDNN
// kernels it replaces) and the residual and the biases are added in its epilogue: 3 reads + 1 write.
//
// CTA tile: 128 output channels (A operand = the weight tile, TMEM lanes) x 256 rows m (B operand = the activation tile,
// TMEM columns).  The K loop walks x then x2 (two tensor maps) in 32-k stages through a 2-stage ring.  The residual tile
// h2[256 rows][128 channels] (128 KiB) is requested by TMA when the CTA starts, so its bytes are in flight during the whole
// main loop; the epilogue adds the accumulators and the bias INTO that shared-memory tile (a lane owns a channel: 32 lanes =
// 128 consecutive bytes, conflict-free) and one bulk tensor store writes it out.  (A first version read h2 and wrote out
// with 4-byte accesses from the epilogue's registers: 32 KiB in flight per SM, latency-bound at the speed of the three
// kernels it replaced.)  The kernel is bound by HBM: 536 MB per 64^2 block against 17 GFLOP.
// CTA: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue (one TMEM lane quarter each).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "umma.cuh"

namespace bndm {

constexpr int kScThreads = 192;
constexpr int kScStages = 2;
constexpr int kScRows = 256;                              // rows m per CTA tile
constexpr uint32_t kScWTile = kBlk * kStageK * 4;         // 16 KiB: 128 channels x 32 k
constexpr uint32_t kScXTile = kScRows * kStageK * 4;      // 32 KiB: 256 rows x 32 k
constexpr uint32_t kScStageBytes = kScWTile + kScXTile;
constexpr uint32_t kScOutTile = kScRows * kBlk * 4;       // 128 KiB: 256 rows x 128 channels, row-major, no swizzle
constexpr uint32_t kScTmemCols = 256;

struct ScArgs {
  const float *bias;
  int M, N;
  int n_nt;                 // output-channel tiles (blockIdx.x % n_nt)
  int ks1, ks2;             // k-stages held by x and by x2
};

// 2-D tiled TMA store shared -> global (rows / columns past the tensor's end are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(kScThreads, 1)
shortcut_tc_kernel(const ScArgs a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_h2,
                   const __grid_constant__ CUtensorMap map_out) {
  extern __shared__ __align__(1024) uint8_t sc_smem[];
  uint8_t *base = sc_smem + ((1024u - (smem_u32(sc_smem) & 1023u)) & 1023u);      // keeps the shared address space
  float *tile = reinterpret_cast<float *>(base + (size_t)kScStages * kScStageBytes);        // the residual / output tile
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + (size_t)kScStages * kScStageBytes + kScOutTile);
  uint64_t *empty_bar = full_bar + kScStages;
  uint64_t *acc_full = empty_bar + kScStages;
  uint64_t *h2_full = acc_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(h2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % a.n_nt, mt = blockIdx.x / a.n_nt;
  const int n_stages = a.ks1 + a.ks2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    if (a.ks2) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    for (int s = 0; s < kScStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(h2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kScTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer: the residual tile first, then a weight tile and an activation tile per stage =================
    if (elect_one()) {
      mbar_expect_tx(h2_full, kScOutTile);
      tma_load_2d(smem_u32(tile), &map_h2, h2_full, nt * kBlk, mt * kScRows, kEvictFirst);
    }
    __syncwarp();
    for (int it = 0; it < n_stages; ++it) {
      const int st = it % kScStages;
      const uint32_t ph = (uint32_t)(it / kScStages) & 1u;
      mbar_wait(&empty_bar[st], ph ^ 1u);
      const uint32_t sa = smem_u32(base + (size_t)st * kScStageBytes);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[st], kScStageBytes);
        tma_load_2d(sa, &map_w, &full_bar[st], it * kStageK, nt * kBlk, kEvictLast);          // weights: read by every row tile
        if (it < a.ks1) tma_load_2d(sa + kScWTile, &map_x, &full_bar[st], it * kStageK, mt * kScRows, kEvictFirst);
        else tma_load_2d(sa + kScWTile, &map_x2, &full_bar[st], (it - a.ks1) * kStageK, mt * kScRows, kEvictFirst);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer: D[128 channels][256 rows] += w tile x activation tile^T =================
    constexpr uint32_t idesc = umma_idesc(kScRows);
    for (int it = 0; it < n_stages; ++it) {
      const int st = it % kScStages;
      const uint32_t ph = (uint32_t)(it / kScStages) & 1u;
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(base + (size_t)st * kScStageBytes);
      if (elect_one()) {
        const uint64_t dW = umma_desc(sa), dX = umma_desc(sa + kScWTile);
#pragma unroll
        for (int kk = 0; kk < kStageK / kUmmaK; ++kk) {
          const uint64_t adv = (uint64_t)((kk * kUmmaK * 4) >> 4);        // +32 B inside the swizzle row
          umma_tf32(tmem_base, dW + adv, dX + adv, idesc, (it | kk) != 0);
        }
        umma_commit(&empty_bar[st]);        // frees the stage when these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else {
    // ================= epilogue: tile[m][n] += D[n][m] + bias[n] in shared memory, then one bulk tensor store =================
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int nl = q * 32 + lane;         // channel inside the tile
    const int n = nt * kBlk + nl;
    const float bias = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.0f;
    mbar_wait(h2_full, 0u);
    mbar_wait(acc_full, 0u);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float *col = tile + nl;
    for (int j = 0; j < kScRows; j += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)j, v);
      float r[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = col[(j + e) * kBlk];
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) col[(j + e) * kBlk] = __fadd_rn(__fadd_rn(__uint_as_float(v[e]), bias), r[e]);
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the bulk store
    asm volatile("bar.sync 1, 128;" ::: "memory");                     // the four epilogue warps
    if (threadIdx.x == 64) {
      tma_store_2d(&map_out, smem_u32(tile), nt * kBlk, mt * kScRows);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile has been read: the CTA may leave
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kScTmemCols);
  }
}

// [rows][K] fp32 row-major, box = box_rows x 32 k, 128-byte swizzle, rows past the end read as zero
static bool sc_map(CUtensorMap *m, const float *ptr, long long rows, int K, int box_rows) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)kStageK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows][N] fp32 row-major, box = 256 rows x 128 channels, no swizzle (the epilogue's tile); edges zero-filled / clipped
static bool sc_tile_map(CUtensorMap *m, const float *ptr, long long rows, int N) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBlk, (cuuint32_t)kScRows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// x: [M][C1], x2: [M][C2] or null (C2 = 0), w: [N][C1 + C2], h2: [M][N], bias: [N] or null, out: [M][N] (may alias h2)
cudaError_t launch_shortcut_tc(const float *x, const float *x2, int C1, int C2, const float *w, const float *h2, const float *bias,
                               float *out, long long M, int N, cudaStream_t s) {
  if (M < 1 || N < 4 || N % 4 != 0 || C1 < kStageK || C1 % kStageK != 0 || C2 < 0 || C2 % kStageK != 0 || (C2 > 0 && !x2)) return cudaErrorNotSupported;
  if (M > 0x7fffffffLL - kScRows) return cudaErrorNotSupported;
  CUtensorMap map_w, map_x, map_x2, map_h2, map_out;
  memset(&map_x2, 0, sizeof(map_x2));
  if (!sc_map(&map_w, w, N, C1 + C2, kBlk) || !sc_map(&map_x, x, M, C1, kScRows) || (C2 > 0 && !sc_map(&map_x2, x2, M, C2, kScRows)) ||
      !sc_tile_map(&map_h2, h2, M, N) || !sc_tile_map(&map_out, out, M, N)) {
    set_error("shortcut_tc: cuTensorMapEncodeTiled failed (M %lld, N %d, C1 %d, C2 %d)", M, N, C1, C2);
    return cudaErrorInvalidValue;
  }
  ScArgs g;
  g.bias = bias; g.M = (int)M; g.N = N;
  g.n_nt = (N + kBlk - 1) / kBlk;
  g.ks1 = C1 / kStageK; g.ks2 = C2 / kStageK;
  const long long n_mt = (M + kScRows - 1) / kScRows;
  const size_t smem = (size_t)kScStages * kScStageBytes + kScOutTile + 1024 + (2 * kScStages + 2) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(shortcut_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  shortcut_tc_kernel<<<(unsigned)(g.n_nt * n_mt), kScThreads, smem, s>>>(g, map_w, map_x, map_x2, map_h2, map_out);
  return cudaGetLastError();
}

}  // namespace bndm
