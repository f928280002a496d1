// K1b (tensor-core variant): the triangular contraction  bn[j][p] = sum_k L[p][k] * z[j][k]
// on tcgen05 with TMA-staged operands and TMEM accumulators.  sm_100a only.
//
// PERSISTENT, STREAM-K.  The work is the list of "k-stages" (one 128-row tile of L x 32 k
// values x one column block) of every row tile, triangular rows first to last:
//     row tile i needs k < 128 (i + 1)  =>  4 (i + 1) stages   (lower-triangular L)
//                                          128 stages           (dense L)
// The W stages are cut into G = min(#SMs, W) equal contiguous ranges, one per CTA, so every
// SM streams the same number of L bytes and issues the same number of MMAs (+-1 stage),
// whatever the triangle looks like.  A CTA's range crosses row-tile borders; each maximal
// piece inside one row tile is a SEGMENT that accumulates in its own TMEM buffer and is
// written out as one partial tile  P[slot][column][128 rows], slot = cta + tile index
// (strictly increasing along the global order, hence unique).  The combine kernel
// (noise_epilogue.cu) adds the partials of a row tile in ascending k order in fp32 --
// deterministic, no atomics.
//
// fp32-grade accuracy from TF32 tensor cores by error compensation (3xTF32):
//   L = Lh + Ll, z = zh + zl (each part exactly representable in tf32, hi rounded rna)
//   L*z ~= Ll*zh + Lh*zl + Lh*zh          (the dropped Ll*zl term is ~2^-22 relative)
//
// CTA = 6 warps: warp 0 = TMA producer (1 lane), warp 1 = TMEM alloc + MMA issuer (1 lane),
// warps 2..5 = epilogue (TMEM -> registers -> partial tile).  Three pipelines: the smem
// stage ring (TMA <-> MMA), two TMEM accumulator buffers (MMA <-> epilogue, so a segment's
// drain overlaps the next segment's MMAs), and the static stream-K schedule.
// Operand tiles are [rows][32 fp32] = 128-byte rows in the canonical K-major SWIZZLE_128B
// layout that both TMA (CU_TENSOR_MAP_SWIZZLE_128B) and the UMMA smem descriptor expect.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace bndm {

constexpr int kATileBytes = kBlk * kStageK * 4;   // 16 KiB
constexpr int kUmmaK = 8;                         // tf32: 32 bytes of K per tcgen05.mma
constexpr int kThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 27;         // bounded spins: a protocol bug traps instead of hanging the GPU

// L2 eviction-priority policies for TMA loads (createpolicy encodings)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // L: streamed once per column block
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;    // z: re-read by every row tile

// ---------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = thread)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile start 1024-byte aligned;
// 8-row groups are 1024 bytes apart).  Bits: [0,14) addr>>4 | [16,30) LBO>>4 (unused for
// swizzled K-major, 1) | [32,46) SBO>>4 | [46,48) version=1 (sm_100) | [61,64) layout=2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ------------------------------------------------------------------------------- kernel
struct TcKernelArgs {
  float *partials;
  int nb;             // columns per column block = UMMA N
  int stages;         // smem ring depth
  uint32_t tmem_cols; // 2 accumulator buffers of buf_cols columns (power of two)
  uint32_t buf_cols;  // power of two >= max(32, nb)
  uint32_t idesc;
  StreamK sk;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_Lh, const __grid_constant__ CUtensorMap map_Ll,
               const __grid_constant__ CUtensorMap map_zh, const __grid_constant__ CUtensorMap map_zl,
               const TcKernelArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages] x {A_hi, A_lo, B_hi, B_lo}, then barriers
  const uint32_t b_tile_bytes = (uint32_t)a.nb * kStageK * 4;
  const uint32_t stage_bytes = 2 * kATileBytes + 2 * b_tile_bytes;
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + (size_t)a.stages * stage_bytes);
  uint64_t *empty_bar = full_bar + a.stages;
  uint64_t *acc_full = empty_bar + a.stages;      // [2]
  uint64_t *acc_empty = acc_full + 2;             // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const StreamK sk = a.sk;
  const int cta = blockIdx.x;
  const int g_begin = sk.cta_begin(cta), g_end = sk.cta_begin(cta + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_Lh);
    prefetch_tmap(&map_Ll);
    prefetch_tmap(&map_zh);
    prefetch_tmap(&map_zl);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);                // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int it = 0;
      for (int g = g_begin; g < g_end;) {
        int cb, tile, s0;
        sk.decode(g, cb, tile, s0);
        const int seg_end = min(g_end, sk.tile_end(cb, tile));
        const int row0 = tile * kBlk, col0 = cb * a.nb;
        for (int s = s0; s < s0 + (seg_end - g); ++s, ++it) {
          const int st = it % a.stages;
          const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
          mbar_wait(&empty_bar[st], ph ^ 1u);
          const uint32_t sa = smem_u32(base + (size_t)st * stage_bytes);
          const int k0 = s * kStageK;
          mbar_expect_tx(&full_bar[st], stage_bytes);
          tma_load_2d(sa, &map_Lh, &full_bar[st], k0, row0, kEvictFirst);
          tma_load_2d(sa + kATileBytes, &map_Ll, &full_bar[st], k0, row0, kEvictFirst);
          tma_load_2d(sa + 2 * kATileBytes, &map_zh, &full_bar[st], k0, col0, kEvictLast);
          tma_load_2d(sa + 2 * kATileBytes + b_tile_bytes, &map_zl, &full_bar[st], k0, col0, kEvictLast);
        }
        g = seg_end;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int it = 0, seg = 0;
      for (int g = g_begin; g < g_end; ++seg) {
        int cb, tile, s0;
        sk.decode(g, cb, tile, s0);
        const int seg_end = min(g_end, sk.tile_end(cb, tile));
        const int buf = seg & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)seg >> 1) & 1u) ^ 1u);     // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * a.buf_cols;
        for (int n = 0; n < seg_end - g; ++n, ++it) {
          const int st = it % a.stages;
          const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(base + (size_t)st * stage_bytes);
          const uint64_t dAh = umma_desc(sa);
          const uint64_t dAl = umma_desc(sa + kATileBytes);
          const uint64_t dBh = umma_desc(sa + 2 * kATileBytes);
          const uint64_t dBl = umma_desc(sa + 2 * kATileBytes + b_tile_bytes);
#pragma unroll
          for (int kk = 0; kk < kStageK / kUmmaK; ++kk) {
            const uint64_t adv = (uint64_t)((kk * kUmmaK * 4) >> 4);   // +32 B inside the swizzle row
            umma_tf32(tmem_d, dAl + adv, dBh + adv, a.idesc, (n | kk) != 0);
            umma_tf32(tmem_d, dAh + adv, dBl + adv, a.idesc, 1u);
            umma_tf32(tmem_d, dAh + adv, dBh + adv, a.idesc, 1u);
          }
          umma_commit(&empty_bar[st]);       // frees the smem stage when these MMAs retire
        }
        umma_commit(&acc_full[buf]);         // this segment's accumulator is complete
        g = seg_end;
      }
    }
  } else {
    // ================= epilogue: TMEM -> partial tile =================
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;          // row inside the tile
    int seg = 0;
    for (int g = g_begin; g < g_end; ++seg) {
      int cb, tile, s0;
      sk.decode(g, cb, tile, s0);
      const int seg_end = min(g_end, sk.tile_end(cb, tile));
      const int buf = seg & 1;
      mbar_wait(&acc_full[buf], ((uint32_t)seg >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * a.buf_cols;
      float *P = a.partials + ((int64_t)sk.slot(cta, cb, tile) * a.nb) * kBlk + r;
      int c = 0;
      for (; c + 32 <= a.nb; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) P[(int64_t)(c + e) * kBlk] = __uint_as_float(v[e]);   // 32 lanes -> 128 B rows
      }
      if (c < a.nb) {                     // nb is a multiple of 16
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) P[(int64_t)(c + e) * kBlk] = __uint_as_float(v[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      g = seg_end;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// --------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D fp32 row-major [rows][4096] tensor, box = [box_rows][32], 128-byte swizzle
static bool make_map(CUtensorMap *m, const float *ptr, int rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)kNPix, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kNPix * 4};
  cuuint32_t box[2] = {(cuuint32_t)kStageK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int tc_max_nb() {
  static int v = 0;
  if (!v) {
    v = 256;
    if (const char *e = getenv("BNDM_TC_MAX_NB")) {
      const int x = atoi(e);
      if (x >= 16 && x <= 256) v = x / 16 * 16;
    }
  }
  return v;
}

int tc_pick_nb(int n_cols) {
  const int cap = tc_max_nb();
  const int pad16 = (n_cols + 15) / 16 * 16;
  if (pad16 <= cap) return pad16;
  // several column blocks: the largest block <= cap that keeps padding small
  const int blocks = (pad16 + cap - 1) / cap;
  const int per = (pad16 + blocks - 1) / blocks;
  return (per + 15) / 16 * 16;
}

int tc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
      n = 148;
  }
  return n;
}

cudaError_t launch_gemm_tc(const TcGemmArgs &g, cudaStream_t s) {
  CUtensorMap mLh, mLl, mzh, mzl;
  if (!make_map(&mLh, g.L_hi, kNPix, kBlk) || !make_map(&mLl, g.L_lo, kNPix, kBlk) ||
      !make_map(&mzh, g.z_hi, g.n_cols_pad, g.nb) || !make_map(&mzl, g.z_lo, g.n_cols_pad, g.nb)) {
    set_error("cuTensorMapEncodeTiled failed (nb=%d, cols=%d)", g.nb, g.n_cols_pad);
    return cudaErrorInvalidValue;
  }
  TcKernelArgs a;
  a.partials = g.partials;
  a.nb = g.nb;
  a.sk = g.sk;
  const uint32_t stage_bytes = 2 * kATileBytes + 2 * (uint32_t)g.nb * kStageK * 4;
  int stages = (int)((224u * 1024u) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  a.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)g.nb) cols <<= 1;
  a.buf_cols = cols;
  a.tmem_cols = 2 * cols;
  // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major A/B,
  // N>>3 at bit 17, M>>4 at bit 24
  a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.nb >> 3) << 17) | ((uint32_t)(kBlk >> 4) << 24);

  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + (2 * stages + 4) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  if (e != cudaSuccess) return e;
  gemm_tc_kernel<<<g.sk.G, kThreads, smem, s>>>(mLh, mLl, mzh, mzl, a);
  return cudaGetLastError();
}

}  // namespace bndm
