// K1b (tensor-core variant): the triangular contraction  bn[j][p] = sum_k L[p][k] * z[j][k]
// on tcgen05 with TMA-staged operands and TMEM accumulators.  sm_100a only.
//
// OPERAND LAYOUT.  A "k-stage" is one 128-row tile of L x 32 k values.  L lives in global memory
// as a sequence of stage blocks that already have the shared-memory image the tensor core wants
// (K-major, 128-byte rows, SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7) of
// its 8-row / 1024-byte group), so ONE linear cp.async.bulk brings a block (or two adjacent
// ones) in -- no tensor map, full-burst DRAM reads.  Only the blocks a lower-triangular L needs
// exist: row tile i has 4 (i + 1) of them, starting at block 2 i (i + 1).  Two variants:
//   pre-split (kRawL = false)  L block = [ Lh tile | Ll tile ] (32 KiB, hi rounded to nearest);
//                              z block (col block cb, k-stage s) = [ zh rows | zl rows ], written
//                              by pack_kernel in the same swizzled image, one more bulk copy
//   raw       (kRawL = true)   L block = the fp32 tile itself (16 KiB); z is the caller's
//                              [columns][4096] tensor read through a 2-D tensor map (hardware
//                              swizzle, rows past n_cols zero-filled) -- no pack kernel; four
//                              converter warps write lo = v - trunc_tf32(v) next to the raw tiles
//                              in shared memory (the tensor core ignores the 13 low mantissa
//                              bits of an fp32 operand, so the raw tile IS the hi operand)
//
// fp32-GRADE ACCURACY FROM TF32 TENSOR CORES.  L = Lh + Ll, z = zh + zl, and
// L z ~= Lh zh + (Lh zl + Ll zh); the dropped Ll zl term is ~2^-22 relative.  Per 8 k values the
// issuer sends TWO MMAs:  Lh x [zh | zl]  (N = 2 nb: main product and first correction side by
// side) and  Ll x zh  (N = nb: second correction), so Lh is read from shared memory once and the
// large main sums never share an accumulator with the small corrections.  The tensor core ADDS
// INTO TMEM WITH TRUNCATION (measured: error grows with the length of the in-TMEM chain and is
// biased towards zero), so a chain is cut after `chain` k-stages (4 => 16 adds): the epilogue
// warps pull the accumulators out of TMEM, add them in fp32 round-to-nearest into per-thread
// running sums, and the issuer carries on in the other TMEM buffer meanwhile.
//
// PERSISTENT, STREAM-K.  The W = (column blocks) x (schedule units) are cut into G = min(#SMs, W)
// equal contiguous ranges, one per CTA (row tiles visited longest first), so every SM streams the
// same number of L bytes whatever the triangle looks like.  A CTA's range crosses row-tile
// borders; each maximal piece inside one row tile is a SEGMENT written out as one partial tile
// P[slot][column][128 rows], slot = cta + visit index (unique).  The combine kernel
// (noise_epilogue.cu) adds the partials of a row tile in ascending k order in fp32 --
// deterministic, no atomics on data.  (Optionally the last CTA to finish a row tile combines it
// in-kernel: implemented, tested, off by default -- see DESIGN.md.)
//
// CTA: warp 0 = producer, warp 1 = TMEM alloc + MMA issuer (both warp-uniform with elect.sync, so
// descriptors stay in uniform registers), warps 2..5 = epilogue, warps 6..9 = converter (raw
// variant).  Pipelines: smem stage ring (producer <-> [converter <->] issuer), two TMEM
// accumulator buffers (issuer <-> epilogue), the static stream-K schedule.  Launched with
// programmatic stream serialisation: L loads go out before griddepcontrol.wait.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "out_map.cuh"
#include "ptx.cuh"
#include "umma.cuh"

namespace bndm {

constexpr int kThreads = 192;
constexpr int kThreadsRaw = 320;                  // + 4 converter warps
// ------------------------------------------------------------------------------- kernel
struct TcKernelArgs {
  const float *Lt;    // stage blocks of L (hi | lo)
  const float *zt;    // stage blocks of z (hi rows | lo rows)
  float *partials;    // [n_slots][NB][128]
  int stages;         // smem ring depth
  int chain;          // stages per TMEM accumulation chain
  uint64_t policy_L;
  unsigned long long *trace;   // debug: [cta][24] time stamps / cycle sums (null in production)
  StreamK sk;
  // fused combine (out != null): the CTA that completes a row tile's last segment sums the
  // tile's partial tiles in ascending-k order and writes the outputs
  int *tile_counters;
  OutMap om;
  int n_cols;
};


// One (<= chain)-stage piece of a segment; every role walks the same sequence.
struct ChainWalk {
  const StreamK sk;
  const int g_end, chain;
  int g;
  int cb, tile, s0, n;      // current chain: stages [s0, s0 + n) of row tile `tile`, column block cb
  bool first, last;         // first / last chain of its segment
  int seg_end;
  __device__ ChainWalk(const StreamK &k, int cta, int chain_) : sk(k), g_end(k.cta_begin(cta + 1)), chain(chain_) {
    g = k.cta_begin(cta);
    seg_end = g;
  }
  __device__ bool next() {
    if (g >= g_end) return false;
    first = (g == seg_end);
    if (first) {
      sk.decode(g, cb, tile, s0);
      seg_end = min(g_end, sk.tile_end(cb, tile));
    } else {
      s0 += n;
    }
    n = min(chain, seg_end - g);
    g += n;
    last = (g == seg_end);
    return true;
  }
};

// kRawL: the L stage blocks in HBM hold the raw fp32 tile only (16 KiB instead of 32): the tensor
// core ignores the 13 low mantissa bits of an fp32 operand (measured, tools/exp_split.py), so the
// raw tile IS the hi operand, and four converter warps (6..9) write  Ll = L - trunc_tf32(L)
// (exact in fp32) next to it in shared memory before the issuer may read the stage.  Halves the
// bytes streamed from HBM; used when one column block covers the call (HBM-bound regime).
// The raw variant also takes z as the caller's fp32 columns [n_cols][4096] through a 2-D tensor
// map (hardware swizzle, rows past n_cols zero-filled) and converts it the same way, so a call
// whose white field is already in column order runs NO pack kernel.  Its accumulators are laid
// out as kSets x [main | Lh.zl | Ll.zh]: consecutive MMAs never target the same TMEM columns
// (back-to-back MMAs into one accumulator serialise on the tensor pipe's latency when N is small).
// kSub = 2: a pipeline stage holds TWO consecutive k-stages (64 k): their raw L tiles are adjacent
// in HBM and come in with one 32 KiB bulk copy -- at 16 KiB per request the memory system
// delivered ~4 TB/s to this kernel, at 32 KiB ~7 (measured) -- and barrier round trips, the
// converter's fence and the commit are paid once per 64 k.  Shared-memory image of a stage:
//   [ L hi/raw x kSub | L lo x kSub | (z hi/raw | z lo) x kSub ]
template <int NB, bool kRawL, int kSub>
__global__ void __launch_bounds__(kRawL ? kThreadsRaw : kThreads, 1)
gemm_tc_kernel(const TcKernelArgs a, const __grid_constant__ CUtensorMap map_z) {
  constexpr uint32_t kTileBytes = kBlk * kStageK * 4;             // 16 KiB: one 128 x 32 fp32 operand tile
  constexpr uint32_t kLBlockBytes = 2 * kTileBytes;               // hi + lo tile of one k-stage in smem
  constexpr uint32_t kLLoadBytes = kRawL ? kTileBytes : kLBlockBytes;           // bytes per k-stage from HBM
  constexpr uint32_t kZBlockBytes = 2 * NB * kStageK * 4;         // zh rows | zl rows
  constexpr uint32_t kZLoadBytes = kRawL ? kZBlockBytes / 2 : kZBlockBytes;
  constexpr uint32_t kStageBytes = kSub * (kLBlockBytes + kZBlockBytes);
  constexpr uint32_t kLoOff = kSub * kTileBytes;                  // lo tiles start here
  constexpr uint32_t kZOff = kSub * kLBlockBytes;                 // z blocks start here
  static_assert(kSub == 1 || kRawL, "two k-stages per pipeline stage only in the raw-operand variant");
  constexpr int kSets = kRawL ? (NB <= 32 ? 2 : 1) : 1;           // independent accumulator sets (alternate per 8 k)
  constexpr uint32_t kSetCols = kRawL ? 3 * NB : 2 * NB;          // raw: main | c1 | c2;  pre-split: main | c1+c2
  constexpr uint32_t kBufCols = kSets * kSetCols;
  constexpr uint32_t kTmemNeed = 2 * kBufCols;
  constexpr uint32_t kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "accumulators do not fit in TMEM");
  static_assert(NB % 16 == 0 && NB >= 16 && NB <= 128, "column block must be a multiple of 16 in [16, 128]");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic)
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + (size_t)a.stages * kStageBytes);
  uint64_t *empty_bar = full_bar + a.stages;
  uint64_t *acc_full = empty_bar + a.stages;      // [2]
  uint64_t *acc_empty = acc_full + 2;             // [2]
  uint64_t *conv_bar = acc_empty + 2;             // [stages] (kRawL): Ll tile written, stage ready for the issuer
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(conv_bar + a.stages);
  volatile int *combine_flag = reinterpret_cast<volatile int *>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  unsigned long long *tr = a.trace ? a.trace + (size_t)cta * 24 : nullptr;
  if (tr && threadIdx.x == 0) { tr[0] = gtime(); tr[1] = clock64(); }

  pdl_launch_dependents();                // the combine kernel may be scheduled; it waits for this grid
  if (warp == 0 && lane > 0 && cta < 2 && a.om.out != nullptr && a.om.gamma != nullptr) {
    // fused combine reads gamma at the very end, when the DRAM queues are full of L traffic and a
    // miss costs several microseconds (measured): pull its few lines into L2 now (hint only)
    for (int i = (lane - 1) * 32; i < a.om.B; i += 31 * 32)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.om.gamma + i));
  }
  if (kRawL && warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], 4);                 // one arrival per converter warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);                // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  // Set-up barrier (named barrier 2): the producer warp only ARRIVES -- its first loads go out
  // while warp 1 is still allocating TMEM (it never needs the TMEM address); everyone else waits.
  constexpr int kAllThreads = kRawL ? kThreadsRaw : kThreads;
  uint32_t tmem_base = 0;
  if (warp == 0) {
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"n"(kAllThreads) : "memory");
  } else {
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(kAllThreads) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }
  if (tr && threadIdx.x == 0) tr[2] = clock64();

  if (warp == 0) {
    // ================= producer: one bulk copy per operand per stage =================
    {
      // L is never written by a preceding kernel: its loads go out at once; the z blocks are
      // the pack kernel's output, so the first of them waits for that grid (PDL)
      bool z_ready = false;
      ChainWalk w(a.sk, cta, a.chain);
      int it = 0;
      while (w.next()) {
        // w.s0 / w.n count schedule units of kSub k-stages
        const float *Lblk = a.Lt + (size_t)(a.sk.cum(w.tile) + w.s0) * kSub * (kLLoadBytes / 4);
        const float *zblk = a.zt + (size_t)(w.cb * (kNPix / kStageK) + w.s0 * kSub) * (kZBlockBytes / 4);
        for (int n = 0; n < w.n; ++n, ++it) {
          const int st = it % a.stages;
          const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
          const long long t0 = tr ? clock64() : 0;
          mbar_wait(&empty_bar[st], ph ^ 1u);
          if (tr && lane == 0) tr[20] += (unsigned long long)(clock64() - t0);
          const uint32_t sa = smem_u32(base + (size_t)st * kStageBytes);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[st], kSub * (kLLoadBytes + kZLoadBytes));
            bulk_load(sa, Lblk + (size_t)n * kSub * (kLLoadBytes / 4), kSub * kLLoadBytes, &full_bar[st], a.policy_L);
          }
          if (!z_ready) {
            pdl_wait();
            z_ready = true;
          }
          if (elect_one()) {
            if (kRawL) {
#pragma unroll
              for (int sb = 0; sb < kSub; ++sb)
                tma_load_2d(sa + kZOff + sb * kZBlockBytes, &map_z, &full_bar[st], ((w.s0 + n) * kSub + sb) * kStageK, w.cb * NB,
                            kEvictLast);
            } else {
              bulk_load(sa + kZOff, zblk + (size_t)n * (kZBlockBytes / 4), kZBlockBytes, &full_bar[st], kEvictLast);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    {
      constexpr uint32_t idesc_wide = umma_idesc(2 * NB), idesc_narrow = umma_idesc(NB);
      ChainWalk w(a.sk, cta, a.chain);
      int it = 0, ch = 0;
      for (; w.next(); ++ch) {
        const int buf = ch & 1;
        const long long ta = tr ? clock64() : 0;
        mbar_wait(&acc_empty[buf], (((uint32_t)ch >> 1) & 1u) ^ 1u);      // epilogue drained this buffer
        tc_fence_after();
        if (tr && lane == 0) tr[19] += (unsigned long long)(clock64() - ta);
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * kBufCols;
        for (int n = 0; n < w.n; ++n, ++it) {
          const int st = it % a.stages;
          const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
          const long long t0 = tr ? clock64() : 0;
          mbar_wait(kRawL ? &conv_bar[st] : &full_bar[st], ph);
          tc_fence_after();
          if (tr && lane == 0) tr[18] += (unsigned long long)(clock64() - t0);
          if (tr && it == 0 && lane == 0) tr[3] = clock64();
          const uint32_t sa = smem_u32(base + (size_t)st * kStageBytes);
          if (elect_one()) {
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            const uint64_t dAh = umma_desc(sa + sb * kTileBytes);
            const uint64_t dAl = umma_desc(sa + kLoOff + sb * kTileBytes);
            const uint64_t dB = umma_desc(sa + kZOff + sb * kZBlockBytes);
#pragma unroll
            for (int kk = 0; kk < kStageK / kUmmaK; ++kk) {
              const uint64_t adv = (uint64_t)((kk * kUmmaK * 4) >> 4);   // +32 B inside the swizzle row
              if (kRawL) {
                const uint32_t d = tmem_d + (uint32_t)(kk % kSets) * kSetCols;
                const uint32_t fresh = (n == 0 && sb == 0 && kk < kSets) ? 0u : 1u;    // first MMA into this set
                umma_tf32(d, dAh + adv, dB + adv, idesc_wide, fresh);                  // Lh x [zh | zl] -> main | c1
                umma_tf32(d + 2 * NB, dAl + adv, dB + adv, idesc_narrow, fresh);       // Ll x zh        -> c2
              } else {
                umma_tf32(tmem_d, dAh + adv, dB + adv, idesc_wide, (n | kk) != 0);     // Lh x [zh | zl]
                umma_tf32(tmem_d + NB, dAl + adv, dB + adv, idesc_narrow, 1u);         // Ll x zh
              }
            }
          }
          umma_commit(&empty_bar[st]);       // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (tr && lane == 0) tr[21] += (unsigned long long)(clock64() - t0);
        }
        if (elect_one()) umma_commit(&acc_full[buf]);         // this chain's accumulators are complete
        __syncwarp();
      }
      if (tr && lane == 0) tr[4] = clock64();
    }
  } else if (kRawL && warp >= 6) {
    // ================= converter: Ll = L - trunc_tf32(L), 8 x 16 bytes per thread per stage =================
    const int t = threadIdx.x - 192;
    ChainWalk w(a.sk, cta, a.chain);
    int it = 0;
    while (w.next()) {
      for (int n = 0; n < w.n; ++n, ++it) {
        const int st = it % a.stages;
        const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
        const long long t0 = (tr && t == 0) ? clock64() : 0;
        mbar_wait(&full_bar[st], ph);
        const long long t1 = (tr && t == 0) ? clock64() : 0;
        const float4 *src = reinterpret_cast<const float4 *>(base + (size_t)st * kStageBytes) + t;
        float4 *dst = reinterpret_cast<float4 *>(base + (size_t)st * kStageBytes + kLoOff) + t;
#pragma unroll
        for (int half = 0; half < kSub; ++half) {
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
#pragma unroll
        for (int sb = 0; sb < kSub; ++sb) {  // zl = z - trunc_tf32(z): NB rows x 128 B per k-stage
          const float4 *zs = reinterpret_cast<const float4 *>(base + (size_t)st * kStageBytes + kZOff + sb * kZBlockBytes);
          float4 *zd = reinterpret_cast<float4 *>(base + (size_t)st * kStageBytes + kZOff + sb * kZBlockBytes + kZBlockBytes / 2);
#pragma unroll
          for (int i = t; i < NB * 8; i += 128) {
            const float4 z4 = zs[i];
            float4 lo;
            lo.x = __fsub_rn(z4.x, __uint_as_float(__float_as_uint(z4.x) & 0xFFFFE000u));
            lo.y = __fsub_rn(z4.y, __uint_as_float(__float_as_uint(z4.y) & 0xFFFFE000u));
            lo.z = __fsub_rn(z4.z, __uint_as_float(__float_as_uint(z4.z) & 0xFFFFE000u));
            lo.w = __fsub_rn(z4.w, __uint_as_float(__float_as_uint(z4.w) & 0xFFFFE000u));
            zd[i] = lo;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[st]);
        if (tr && t == 0) { tr[17] += (unsigned long long)(t1 - t0); tr[16] += (unsigned long long)(clock64() - t1); }
      }
    }
  } else {
    // ================= epilogue: TMEM -> fp32 running sums -> partial tile =================
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;          // row inside the tile
    float acc[NB];
    pdl_wait();                           // the previous call's combine is done reading the partial tiles
    ChainWalk w(a.sk, cta, a.chain);
    for (int ch = 0; w.next(); ++ch) {
      const int buf = ch & 1;
      if (w.first) {
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] = 0.0f;
      }
      mbar_wait(&acc_full[buf], ((uint32_t)ch >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kBufCols;
      if (kRawL) {
        // a chain shorter than kSets MMAs per stage cannot happen (4 MMAs per stage, kSets <= 2)
#pragma unroll
        for (int set = 0; set < kSets; ++set)
#pragma unroll
          for (int c = 0; c < NB; c += 16) {
            uint32_t m[16], x[16], y[16];
            const uint32_t t0 = taddr + (uint32_t)set * kSetCols + (uint32_t)c;
            tmem_ld16(t0, m);
            tmem_ld16(t0 + NB, x);
            tmem_ld16(t0 + 2 * NB, y);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e)
              acc[c + e] = __fadd_rn(acc[c + e], __fadd_rn(__uint_as_float(m[e]),
                                                            __fadd_rn(__uint_as_float(x[e]), __uint_as_float(y[e]))));
          }
      } else {
#pragma unroll
        for (int c = 0; c < NB; c += 16) {
          uint32_t m[16], x[16];
          tmem_ld16(taddr + (uint32_t)c, m);
          tmem_ld16(taddr + (uint32_t)(NB + c), x);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            acc[c + e] = __fadd_rn(acc[c + e], __fadd_rn(__uint_as_float(m[e]), __uint_as_float(x[e])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (w.last) {
        const int c_first = a.sk.first_cta(w.cb, w.tile), c_last = a.sk.last_cta(w.cb, w.tile);
        const int col0 = w.cb * NB;
        const int p = w.tile * kBlk + r;
        if (a.om.out != nullptr && c_first == c_last) {
          // the whole row tile was computed here: straight from the registers to the outputs
#pragma unroll
          for (int c = 0; c < NB; c += 16)
            if (col0 + c < a.n_cols) emit_cols<16>(a.om, col0 + c, a.n_cols, p, acc + c);
        } else {
          const int64_t slot_stride = (int64_t)NB * kBlk;
          float *P = a.partials + (int64_t)a.sk.slot(cta, w.cb, w.tile) * slot_stride + r;
#pragma unroll
          for (int c = 0; c < NB; ++c) P[(int64_t)c * kBlk] = acc[c];          // 32 lanes -> 128 B rows
          if (a.om.out != nullptr) {
            // ticket: the last of the tile's (c_last - c_first + 1) contributors combines
            if (tr && threadIdx.x == 64) tr[8] = clock64();
            __threadfence();
            if (tr && threadIdx.x == 64) tr[9] = clock64();
            asm volatile("bar.sync 1, 128;" ::: "memory");                     // the 4 epilogue warps
            if (threadIdx.x == 64) {
              int *cnt = a.tile_counters + w.cb * a.sk.n_tiles + w.tile;
              const int old = atomicAdd(cnt, 1);
              const int is_last = (old == c_last - c_first);
              if (is_last) *cnt = 0;                                           // ready for the next call
              *combine_flag = is_last;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tr && threadIdx.x == 64) tr[10] = clock64();
            if (*combine_flag) {
              __threadfence();
              if (tr && threadIdx.x == 64) { tr[11] = clock64(); tr[14] = (unsigned long long)(c_last - c_first + 1); }
              const float *Q = a.partials + (int64_t)a.sk.slot(c_first, w.cb, w.tile) * slot_stride + r;
              const int n_seg = c_last - c_first + 1;
              // CB columns x SG segments (= 64) independent L2 loads in flight per thread, then the
              // adds in ascending-k order (the order is what makes the result reproducible)
              // (fewer for the big column blocks, whose running sums already fill the register file;
              // they use the wide combine kernel by default)
              constexpr int CB = NB <= 32 ? 16 : 8, SG = NB <= 32 ? 4 : 2;
              for (int c = 0; c < NB; c += CB) {
                if (col0 + c >= a.n_cols) break;
                float v[CB];
#pragma unroll
                for (int e = 0; e < CB; ++e) v[e] = 0.0f;
                for (int s0 = 0; s0 < n_seg; s0 += SG) {
                  float u[SG][CB];
#pragma unroll
                  for (int g = 0; g < SG; ++g)
#pragma unroll
                    for (int e = 0; e < CB; ++e)
                      u[g][e] = (s0 + g < n_seg) ? __ldcg(Q + (int64_t)(s0 + g) * slot_stride + (int64_t)(c + e) * kBlk) : 0.0f;
#pragma unroll
                  for (int g = 0; g < SG; ++g)
                    if (s0 + g < n_seg) {
#pragma unroll
                      for (int e = 0; e < CB; ++e) v[e] = (s0 + g == 0) ? u[g][e] : __fadd_rn(v[e], u[g][e]);
                    }
                }
                if (tr && threadIdx.x == 64) tr[12] = clock64();
                emit_cols<CB>(a.om, col0 + c, a.n_cols, p, v);
              }
              if (tr && threadIdx.x == 64) tr[13] = clock64();
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");                     // combine_flag is reused
          }
        }
      }
    }
    if (tr && threadIdx.x == 64) tr[5] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (tr && threadIdx.x == 0) { tr[6] = clock64(); tr[7] = gtime(); }
}

// --------------------------------------------------------------------------------- host
int tc_max_nb() {
  static int v = 0;
  if (!v) {
    v = 128;
    if (const char *e = getenv("BNDM_TC_MAX_NB")) {
      const int x = atoi(e);
      if (x >= 16 && x <= 128) v = x / 16 * 16;
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
  // per device (a process may drive several GPUs from several threads): a small table indexed by the current device
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;      // benign race: every writer stores the same value
  return n;
}

// Policy knobs: -1 = default rule, 0 = off, 1 = on.  Initialised from BNDM_TC_FUSED / BNDM_TC_RAWL,
// overridable at run time through bndm_debug_set_policy (tests sweep every variant).
static std::atomic<int> g_policy_fused{-2}, g_policy_raw{-2};     // test hooks: process-wide, but race-free
static int env_policy(const char *name) {
  const char *e = getenv(name);
  return e ? (e[0] == '1' ? 1 : 0) : -1;
}
void tc_set_policy(int fused, int raw) {
  g_policy_fused = fused;
  g_policy_raw = raw;
}
bool tc_fused_combine(int nb) {
  if (g_policy_fused == -2) g_policy_fused = env_policy("BNDM_TC_FUSED");
  return g_policy_fused < 0 ? false : g_policy_fused == 1;   // default off: see DESIGN.md (tail latency)
}

static int tc_chain() {
  static int v = 0;
  if (!v) {
    v = 4;
    if (const char *e = getenv("BNDM_TC_CHAIN")) {
      const int x = atoi(e);
      if (x >= 1 && x <= 4096) v = x;
    }
  }
  return v;
}

EncodeTiledFn get_tensormap_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// fp32 columns [rows][4096] row-major, box = [box_rows][32 k], 128-byte swizzle, OOB rows read as zero
static bool make_z_map(CUtensorMap *m, const float *ptr, int rows, int box_rows) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)kNPix, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kNPix * 4};
  cuuint32_t box[2] = {(cuuint32_t)kStageK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NB, bool kRawL, int kSub>
static cudaError_t launch_nb(const TcGemmArgs &g, cudaStream_t s) {
  TcKernelArgs a;
  a.Lt = g.Lt;
  a.zt = g.zt;
  a.partials = g.partials;
  a.sk = g.sk;
  a.chain = tc_chain() / kSub > 0 ? tc_chain() / kSub : 1;       // schedule units per TMEM chain
  a.trace = g.trace;
  a.tile_counters = g.tile_counters;
  a.om = OutMap{g.z_cols, g.gamma, g.out, g.out_bn, g.out_wn, g.B, g.C, g.res_mode, TrainOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
  a.n_cols = g.n_cols;
  // L is streamed once per call when there is one column block; with several, the CTAs of the
  // other column blocks read the same stage blocks at about the same time -> keep them in L2
  a.policy_L = g.sk.n_colblk == 1 ? kEvictFirst : kEvictNormal;
  const uint32_t stage_bytes = kSub * (2 * kBlk * kStageK * 4 + 2 * NB * kStageK * 4);
  const uint32_t budget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;
  // One column block (cold L2): the kernel is HBM-bound and more than ~100 KiB in flight per SM
  // only lengthens the DRAM queue, i.e. spreads the time the first operands land and with it
  // the CTAs' finishing times (measured).  Several column blocks: half the L loads hit in L2 and
  // shared-memory bandwidth bounds a stage -> as deep as fits.
  int stages = (int)(budget / stage_bytes);
  if (stages > 8) stages = 8;
  if (g.sk.n_colblk == 1 && stages > (kRawL && kSub == 1 ? 5 : 3)) stages = kRawL && kSub == 1 ? 5 : 3;
  if (const char *e = getenv("BNDM_TC_STAGES")) {     // experiment knob: shallower ring
    const int x = atoi(e);
    const int fit = (int)(budget / stage_bytes) < 8 ? (int)(budget / stage_bytes) : 8;
    if (x >= 1 && x <= fit) stages = x;
  }
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (3 * stages + 4) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<NB, kRawL, kSub>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  if (e != cudaSuccess) return e;
  CUtensorMap map_z;
  memset(&map_z, 0, sizeof(map_z));
  if (kRawL && !make_z_map(&map_z, g.z_cols, g.n_cols, NB)) {
    set_error("cuTensorMapEncodeTiled failed (columns %d, block %d)", g.n_cols, NB);
    return cudaErrorInvalidValue;
  }
  return launch_pdl(gemm_tc_kernel<NB, kRawL, kSub>, dim3(g.sk.G), dim3(kRawL ? kThreadsRaw : kThreads), smem, s, a, map_z);
}

cudaError_t launch_gemm_tc(const TcGemmArgs &g, cudaStream_t s) {
  if (g.raw_L) {
    if (g.sk.sub == 2) {
      switch (g.nb) {
        case 16: return launch_nb<16, true, 2>(g, s);
        case 32: return launch_nb<32, true, 2>(g, s);
      }
      set_error("tcgen05 contraction (raw L, 64-k stages): unsupported column block %d", g.nb);
      return cudaErrorInvalidValue;
    }
    switch (g.nb) {
      case 16: return launch_nb<16, true, 1>(g, s);
      case 32: return launch_nb<32, true, 1>(g, s);
      case 48: return launch_nb<48, true, 1>(g, s);
      case 64: return launch_nb<64, true, 1>(g, s);
    }
    set_error("tcgen05 contraction (raw L): unsupported column block %d", g.nb);
    return cudaErrorInvalidValue;
  }
  switch (g.nb) {
    case 16: return launch_nb<16, false, 1>(g, s);
    case 32: return launch_nb<32, false, 1>(g, s);
    case 48: return launch_nb<48, false, 1>(g, s);
    case 64: return launch_nb<64, false, 1>(g, s);
    case 80: return launch_nb<80, false, 1>(g, s);
    case 96: return launch_nb<96, false, 1>(g, s);
    case 112: return launch_nb<112, false, 1>(g, s);
    case 128: return launch_nb<128, false, 1>(g, s);
  }
  set_error("tcgen05 contraction: unsupported column block %d", g.nb);
  return cudaErrorInvalidValue;
}

int tc_sub(int nb, bool raw) {
  static int v = -2;           // BNDM_TC_SUB=1/2 forces it (experiments)
  if (v == -2) {
    const char *e = getenv("BNDM_TC_SUB");
    v = e ? atoi(e) : -1;
  }
  if (!raw || nb > 32) return 1;
  if (v == 1 || v == 2) return v;
  return nb == 16 ? 2 : 1;     // 3 stages of 2 x (32 KiB + z) must fit in shared memory
}

bool tc_raw_L(int nb, int n_colblk) {
  if (g_policy_raw == -2) g_policy_raw = env_policy("BNDM_TC_RAWL");
  if (nb > 64) return false;                       // register budget of the 320-thread variant
  return g_policy_raw < 0 ? n_colblk == 1 : g_policy_raw == 1;
}

}  // namespace bndm
