// K1g: get_noise_v2's contraction  bn[j][p] = sum_k L[p][k] * z[j][k]  in the GEMV regime
// (<= 16 GEMM columns, e.g. BASELINE config 1: B=4, C=3 -> 12 columns), sm_100a only.
//
// With so few columns the product is a stream of L (33.5 MB, read exactly once) against a white
// field that fits in L2 forty times over: the bound is HBM, not math (12 FMAs per L element =
// 3.0 us of packed-FFMA issue per call on 148 SMs against >= 5.3 us of DRAM time).  Feeding the
// tensor cores at this shape costs more than it saves: tcgen05 needs 64/128-row tiles, so any
// balanced cut of the triangle needs a split-K reduction pass (K1b + K1c: three launches), and
// 3xTF32 operand splitting moves every L byte through shared memory three more times; the legacy
// mma.sync path runs at 512 MAC/clk/SM on this chip (measured, tools/probes/pipe_probe.cu), no
// faster than packed FFMA once the 3x of the error compensation is paid.  So this kernel keeps
// plain fp32 FMA arithmetic (results are fp32 sums, ~3x closer to the fp64 product than 3xTF32)
// and is organised around the stream:
//
//   * ROW OWNERSHIP, NO SPLIT-K.  The work unit is a QUAD = 4 consecutive rows of L (they end at
//     the same 16-byte column).  The 1024 quads (256 at 32^2, where only h,w < 32 is kept) are
//     dealt to the CTAs longest-first onto the least-loaded CTA (host, once per handle), so every
//     CTA streams the same number of L bytes (+-2 %) and owns COMPLETE rows: no partial tiles in
//     global memory, no combine kernel, one launch per call.
//   * CHUNK-MAJOR STREAM, FOLDED.  A CTA walks k in chunks of 128; the longest of its quads has n
//     chunks.  Pipeline stage s carries chunk s AND chunk n-1-s of every quad that still reaches
//     them: the number of active quads falls linearly with the chunk index, so every stage moves
//     about the same bytes (~8 x 2 KiB of L) and the stream stays bandwidth-bound to its end
//     (chunk by chunk, the late stages carried 2-4 KiB and ran on latency: 2.5 TB/s, measured).
//     A stage = those L blocks (ONE linear cp.async.bulk) + the two chunks of z (TMA boxes NC x 128
//     straight from the caller's tensor, rows past n_cols zero-filled, evict-last: all CTAs read
//     the same 196 KB) behind one mbarrier.  z is never resident, so the column count is not
//     limited by shared memory.
//   * L IS STORED IN STREAM ORDER.  Reading a k-chunk of 28 scattered rows straight from the
//     row-major matrix is a column-stripe access: 512-byte pieces 16 KiB apart (measured: 25 us
//     for the 33.5 MB, DRAM pages opened for a quarter of their bytes).  bndm_prepare_L therefore
//     writes `Lg`, a copy of the lower triangle in exactly the order the kernel consumes it: per
//     CTA, per stage, one 2 KiB block per active quad (a CTA's slots are sorted longest first, so
//     the active ones are a prefix), chunk s first, then chunk n-1-s.  A CTA's whole stream is one
//     contiguous 230 KB region.  Inside a block the rows are interleaved in pairs,
//     [pair][half][lane] x (L[2p][k], L[2p+1][k], L[2p][k+1], L[2p+1][k+1]), so that one LDS.128
//     hands a lane two ready-made operands of fma.rn.f32x2 (rows 2p, 2p+1 in the two halves of one
//     64-bit register) and every shared-memory load of a warp is 512 contiguous bytes.
//   * 16 consumer warps = 8 quad slots x 2 stage parities; a lane owns 4 of the 128 k.  Shared
//     memory bandwidth decides the mapping (measured: an LDS.128 costs 4 cycles whether its 32
//     lanes read 512 distinct bytes or the same 128 bytes four times): z is read once per
//     (quad, chunk) as 12 full-width loads, 48 FMAs per loaded z value.  Sub-partition s hosts
//     slots s and 7-s (long + short), so the four FMA pipes carry equal work and the longest
//     quad's chunks -- the critical path of a CTA -- are spread over two warps.  4 producer warps
//     take the stages round-robin (one thread's wait + expect_tx + copy issues cost ~600 cycles).
//   * The schedule row of a CTA travels in the kernel parameters (constant bank): the first TMA
//     request does not wait for a cold global load.
//   * Deterministic end: lanes dump their sums to shared memory, the consumer threads add the 64
//     partial sums of every output in a fixed order, and the 4 rows of a quad leave as one float4
//     through the same output map as the other paths (lerp, crop, 128^2 placement, training
//     outputs).  The white values / gamma an output needs were fetched before the stream started
//     (parked in shared memory: the accumulators need the registers).  A column's result does not
//     depend on the other columns: K1g is batch-invariant bit for bit.
//
// L above the diagonal is never needed but a quad's last chunk is read whole: that is why the
// caller must have checked triangularity (zeros) -- or pass dense = 1, which walks every chunk.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "out_map.cuh"
#include "ptx.cuh"

namespace bndm {

constexpr int kGvWarps = 16;                        // consumer warps: 8 quad slots x 2 stage parities
constexpr int kGvProducers = 4;                     // producer warps: stage c is requested by warp c % 4 (one thread's
                                                    // wait + expect_tx + two copy issues cost ~600 cycles: measured)
constexpr int kGvThreads = (kGvWarps + kGvProducers) * 32;
constexpr int kGvKW = 128;                          // k extent of a stage
constexpr uint32_t kGvQuadBytes = 4u * kGvKW * 4u;  // one quad's block of a stage: 2 KiB

template <int NC>
struct GvCfg {
  static constexpr uint32_t kZBytes = (uint32_t)NC * kGvKW * 4u;   // one k-chunk of z
  static constexpr int R = 64;                      // partial sums per output: 32 lanes x 2 stage parities
  static constexpr int NOUT = kGemvSlots * 4 * NC;  // outputs per CTA
  static constexpr int PS = NOUT + 1;               // row pitch of the partial-sum dump: odd -> conflict-free both ways
  static constexpr uint32_t kPartBytes = (uint32_t)R * PS * 4u;     // the dump ...
  static constexpr uint32_t kRedBytes = (uint32_t)NOUT * 4u;        // ... and the reduced outputs behind it (16-byte aligned below)
  static_assert(NOUT / 4 <= kGvWarps * 32, "one float4 of outputs per consumer thread");
};

// The CTA's schedule rows travel in the kernel parameters (constant bank): the first TMA request does not wait for a
// cold global load.  kGvMaxCtas bounds the struct (B200: 148 SMs).
constexpr int kGvMaxCtas = 160;
struct GvTable {
  int max_blocks;                                   // most 2 KiB blocks any stage of any CTA carries (sizes the ring's stages)
  int off[kGvMaxCtas];                              // first 2 KiB block of the CTA's stream in Lg
  short quad[kGvMaxCtas][kGemvSlots];               // quad index per slot (-1: empty), longest first
};

struct GemvKernelArgs {
  const float *Lg;     // L in stream order (launch_gemv_pack_L)
  int dense;           // every quad spans all 4096 k
  int stages;          // ring depth
  uint32_t zoff;       // bytes of a stage's L region (max blocks of any stage of any CTA x 2 KiB); z chunks a, b follow
  uint32_t stage_bytes;
  int n_cols;
  OutMap om;
  unsigned long long *trace;   // debug: [cta][24] time stamps (null in production)
};

// d.x += a.x * b.x;  d.y += a.y * b.y   (one FFMA2: both halves of a 64-bit register)
__device__ __forceinline__ void fma2(float2 &d, const float2 a, const float2 b) {
  unsigned long long dd = reinterpret_cast<unsigned long long &>(d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
  d = reinterpret_cast<float2 &>(dd);
}

// Ring slot and mbarrier phase of stage c.  Even stages live in the even slots, odd stages in the odd ones: a consumer
// warp only ever waits for the stages of ITS parity, and a parity wait is only meaningful for a waiter that sees every
// use of the barrier (a warp that skipped one use would take the completion of the use before for the one it waits
// for -- with a plain c % stages ring and an odd depth that happened, rarely, and ended in a trap).  Returns true for
// the first use of the slot.
__device__ __forceinline__ bool gv_ring_slot(int c, int stages, int &slot, uint32_t &phase) {
  const int par = c & 1, k = c >> 1;
  const int n = par ? stages / 2 : (stages + 1) / 2;       // slots of this parity
  const int u = k / n;
  slot = 2 * (k - u * n) + par;
  phase = (uint32_t)u & 1u;
  return u == 0;
}

// one k-chunk of one consumer warp: its quad x NC columns x (this lane's 4 k)
template <int NC>
__device__ __forceinline__ void gv_chunk(const uint8_t *lp, const uint8_t *zp, float2 (&acc)[2][NC]) {
  float4 La[2][2];                                         // [row pair][half]: (L[2p][k], L[2p+1][k], L[2p][k+1], L[2p+1][k+1])
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int h = 0; h < 2; ++h) La[p][h] = *reinterpret_cast<const float4 *>(lp + (p * 2 + h) * 512);
  // two columns per step (4 independent FFMA2 chains), the next two columns' z values already in flight
  float4 zn0 = *reinterpret_cast<const float4 *>(zp), zn1 = *reinterpret_cast<const float4 *>(zp + kGvKW * 4);
#pragma unroll
  for (int j = 0; j < NC; j += 2) {
    const float4 za = zn0, zb = zn1;
    if (j + 2 < NC) {
      zn0 = *reinterpret_cast<const float4 *>(zp + (j + 2) * (kGvKW * 4));
      zn1 = *reinterpret_cast<const float4 *>(zp + (j + 3) * (kGvKW * 4));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float wa = kk == 0 ? za.x : kk == 1 ? za.y : kk == 2 ? za.z : za.w;
      const float wb = kk == 0 ? zb.x : kk == 1 ? zb.y : kk == 2 ? zb.z : zb.w;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float4 lv = La[p][kk >> 1];
        const float2 l2 = (kk & 1) ? make_float2(lv.z, lv.w) : make_float2(lv.x, lv.y);
        fma2(acc[p][j], l2, make_float2(wa, wa));
        fma2(acc[p][j + 1], l2, make_float2(wb, wb));
      }
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(kGvThreads, 1)
gemv_kernel(const GemvKernelArgs a, const __grid_constant__ CUtensorMap map_z, const __grid_constant__ GvTable tab) {
  using Cfg = GvCfg<NC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic)
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + (size_t)a.stages * a.stage_bytes);
  uint64_t *empty_bar = full_bar + a.stages;
  __shared__ OutPos4 s_pos[Cfg::NOUT / 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long *tr = a.trace ? a.trace + (size_t)blockIdx.x * kGemvTraceStride : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = gtime();

  if (threadIdx.x == 32) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kGvWarps / 2);      // one arrival per consumer warp of the stage's parity
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  auto kend_of = [&](int slot) {
    const int Q = tab.quad[blockIdx.x][slot];
    return Q < 0 ? 0 : (a.dense ? kNPix : 4 * Q + 4);
  };
  // FOLDED CHUNK ORDER.  The CTA's longest quad (slot 0) has n k-chunks; stage s carries chunk a = s AND chunk
  // b = n-1-s: the number of active quads falls with the chunk index, so nact(a) + nact(b) -- the bytes of a stage -- is
  // about the same for every stage and the stream stays bandwidth-bound to its end (chunk by chunk the late stages would
  // carry 2-4 KiB each and run on latency: measured 2.5 TB/s there).
  const int n_chunks = (kend_of(0) + kGvKW - 1) / kGvKW;
  const int nst = (n_chunks + 1) / 2;

  if (warp >= kGvWarps) {
    // ================= producers: one linear bulk copy of L + one TMA box of z per stage =================
    const int pw = warp - kGvWarps;
    bool z_ready = false;
    int kend[kGemvSlots];
#pragma unroll
    for (int q = 0; q < kGemvSlots; ++q) kend[q] = kend_of(q);
    const float *Lsrc = a.Lg + (size_t)tab.off[blockIdx.x] * (kGvQuadBytes / 4);
    for (int c = 0; c < nst; ++c) {
      const int ca = c, cb = n_chunks - 1 - c;
      int nact = 0;
#pragma unroll
      for (int q = 0; q < kGemvSlots; ++q) nact += (kend[q] > ca * kGvKW ? 1 : 0) + ((cb > ca && kend[q] > cb * kGvKW) ? 1 : 0);
      if ((c % kGvProducers) == pw) {
        int st;
        uint32_t ph;
        const bool first_use = gv_ring_slot(c, a.stages, st, ph);
        if (!first_use) mbar_wait(&empty_bar[st], ph ^ 1u);          // the first use of a slot finds it free
        const uint32_t sa = smem_u32(base + (size_t)st * a.stage_bytes);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], (uint32_t)nact * kGvQuadBytes + (cb > ca ? 2u : 1u) * Cfg::kZBytes);
          // L is immutable: its load goes out before the grid dependency is resolved.  Chunk a's active slots (a
          // prefix), then chunk b's: one contiguous piece of the CTA's stream.
          bulk_load(sa, Lsrc, (uint32_t)nact * kGvQuadBytes, &full_bar[st], kEvictFirst);
        }
        if (!z_ready) {
          pdl_wait();             // z may be the previous kernel's output (torch.randn, the pack kernel)
          z_ready = true;
        }
        if (elect_one()) {
          tma_load_2d(sa + a.zoff, &map_z, &full_bar[st], ca * kGvKW, 0, kEvictLast);
          if (cb > ca) tma_load_2d(sa + a.zoff + Cfg::kZBytes, &map_z, &full_bar[st], cb * kGvKW, 0, kEvictLast);
        }
        if (tr && lane == 0 && c < 40) tr[8 + c] = gtime();
        __syncwarp();
      }
      Lsrc += (size_t)nact * (kGvQuadBytes / 4);
    }
  } else {
    // ================= consumers: warp = (quad slot, stage parity), lane = 4 of the 128 k =================
    // Sub-partition s hosts slots s (long) and 7-s (short) for both parities: equal FMA work on the four pipes, and the
    // longest quad's chunks are spread over two warps (the critical path of the CTA).
    const int sub = warp & 3, u = warp >> 2;
    const int slot = (u < 2) ? sub : kGemvSlots - 1 - sub;
    const int par = u & 1;
    const int kend = kend_of(slot);
    float2 acc[2][NC];                                                  // [row pair][column] = (row 2p, row 2p+1)
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[p][j] = make_float2(0.f, 0.f);

    pdl_wait();
    // the float4 of outputs this thread will emit at the very end (4 rows of one quad, one column): fetch its white
    // values / gamma / training inputs now, under the stream, and park them in shared memory (not in registers: the
    // accumulators need those)
    bool owner = false;
    if ((int)threadIdx.x < Cfg::NOUT / 4) {
      const int oq = (int)threadIdx.x / NC, oj = (int)threadIdx.x - oq * NC;
      const int Qo = tab.quad[blockIdx.x][oq];
      if (Qo >= 0 && oj < a.n_cols) {
        owner = true;
        s_pos[threadIdx.x] = locate4(a.om, oj, 4 * Qo);
      }
    }

    const uint32_t l_off = (uint32_t)slot * kGvQuadBytes + (uint32_t)lane * 16u;
    const uint32_t z_off = a.zoff + (uint32_t)lane * 16u;
    for (int c = par; c < nst; c += 2) {
      int st;
      uint32_t ph;
      gv_ring_slot(c, a.stages, st, ph);
      const int ca = c, cb = n_chunks - 1 - c;
      int nact_a = 0;                                   // chunk b's blocks sit behind chunk a's active ones
#pragma unroll
      for (int q = 0; q < kGemvSlots; ++q) nact_a += kend_of(q) > ca * kGvKW ? 1 : 0;
      mbar_wait(&full_bar[st], ph);
      if (tr && c == 0 && threadIdx.x == 0) tr[1] = gtime();
      if (tr && slot == 0 && lane == 0 && c < 40) tr[48 + c] = gtime();
      const uint8_t *sp = base + (size_t)st * a.stage_bytes;
      if (kend > ca * kGvKW) gv_chunk<NC>(sp + l_off, sp + z_off, acc);                 // warp-uniform
      if (cb > ca && kend > cb * kGvKW) gv_chunk<NC>(sp + (uint32_t)nact_a * kGvQuadBytes + l_off, sp + z_off + Cfg::kZBytes, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
      if (tr && slot == 0 && lane == 0 && c < 40) tr[88 + c] = gtime();
    }
    if (tr && threadIdx.x == 0) tr[2] = gtime();

    // ---- deterministic reduction of the 64 partial sums of every output, then the output map
    asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");        // every consumer is done with the ring
    float *part = reinterpret_cast<float *>(base);
    float *red = reinterpret_cast<float *>(base + ((Cfg::kPartBytes + 15u) & ~15u));
    {
      float *dst = part + (par * 32 + lane) * Cfg::PS + slot * NC * 4;
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          dst[j * 4 + 2 * p] = acc[p][j].x;
          dst[j * 4 + 2 * p + 1] = acc[p][j].y;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
    for (int o = threadIdx.x; o < Cfg::NOUT; o += kGvWarps * 32) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int r = 0; r < Cfg::R; r += 4) {
        v0 = __fadd_rn(v0, part[(r + 0) * Cfg::PS + o]);
        v1 = __fadd_rn(v1, part[(r + 1) * Cfg::PS + o]);
        v2 = __fadd_rn(v2, part[(r + 2) * Cfg::PS + o]);
        v3 = __fadd_rn(v3, part[(r + 3) * Cfg::PS + o]);
      }
      red[o] = __fadd_rn(__fadd_rn(v0, v1), __fadd_rn(v2, v3));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
    if (owner) store4(a.om, s_pos[threadIdx.x], *reinterpret_cast<const float4 *>(red + 4 * threadIdx.x));
    if (tr && threadIdx.x == 0) tr[3] = gtime();
  }
}

// --------------------------------------------------------------------------------- host
int gemv_variant() { return 0; }     // one instance family today; the handle keeps the number for the stream-order layout
int gemv_kw(int) { return kGvKW; }

bool gemv_policy() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("BNDM_GEMV");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Deals the quads to the CTAs (longest first onto the least-loaded CTA with a free slot); a CTA's slots are sorted
// longest first.  table: [n_ctas][kGemvTableStride]: 8 quad indices (-1: empty), then the index of the CTA's first
// 2 KiB block in the stream-ordered copy of L and its block count.  Returns the total number of blocks, or -1 if the
// quads do not fit.
long gemv_build_schedule(int res32, int dense, int n_ctas, int variant, int *table) {
  (void)variant;
  struct Item { int q, w; };
  std::vector<Item> items;
  for (int q = 0; q < kNPix / 4; ++q) {
    if (res32 && (q >= 512 || (q & 15) >= 8)) continue;          // rows with h >= 32 or w >= 32 are cropped away
    const int kend = dense ? kNPix : 4 * q + 4;
    items.push_back({q, (kend + kGvKW - 1) / kGvKW});
  }
  if ((int)items.size() > n_ctas * kGemvSlots) return -1;
  auto longer = [](const Item &x, const Item &y) { return x.w > y.w || (x.w == y.w && x.q > y.q); };
  std::stable_sort(items.begin(), items.end(), longer);
  std::vector<std::vector<Item>> per(n_ctas);
  std::vector<long> load(n_ctas, 0);
  for (const Item &it : items) {
    int best = -1;
    for (int c = 0; c < n_ctas; ++c)
      if ((int)per[c].size() < kGemvSlots && (best < 0 || load[c] < load[best])) best = c;
    per[best].push_back(it);
    load[best] += it.w;
  }
  long blocks = 0;
  for (int c = 0; c < n_ctas; ++c) {
    std::stable_sort(per[c].begin(), per[c].end(), longer);
    for (int i = 0; i < kGemvSlots; ++i) table[c * kGemvTableStride + i] = i < (int)per[c].size() ? per[c][i].q : -1;
    table[c * kGemvTableStride + kGemvSlots] = (int)blocks;
    table[c * kGemvTableStride + kGemvSlots + 1] = (int)load[c];
    blocks += load[c];
  }
  return blocks;
}

// L -> stream order.  Block (cta, stage c, active slot) covers rows 4Q .. 4Q+3, k in [128 c, 128 c + 128) as
// [pair p][half h][lane l] x float4 (L[4Q+2p][k], L[4Q+2p+1][k], L[4Q+2p][k+1], L[4Q+2p+1][k+1]), k = 128 c + 4 l + 2 h.
__global__ void __launch_bounds__(256) gemv_pack_L_kernel(const float *__restrict__ L, float *__restrict__ Lg,
                                                          const int *__restrict__ table, int dense) {
  const int *t = table + blockIdx.x * kGemvTableStride;
  int kend[kGemvSlots], kmax = 0;
  for (int q = 0; q < kGemvSlots; ++q) {
    kend[q] = t[q] < 0 ? 0 : (dense ? kNPix : 4 * t[q] + 4);
    kmax = max(kmax, kend[q]);
  }
  float4 *dst = reinterpret_cast<float4 *>(Lg + (size_t)t[kGemvSlots] * (kGvQuadBytes / 4));
  const int n_chunks = (kmax + kGvKW - 1) / kGvKW;
  for (int s = 0; 2 * s < n_chunks; ++s) {                       // stage s = chunk s, then chunk n-1-s (folded order)
    for (int half = 0; half < 2; ++half) {
      const int c = half ? n_chunks - 1 - s : s;
      if (half && c <= s) break;
      int nact = 0;
      for (int q = 0; q < kGemvSlots; ++q) nact += kend[q] > c * kGvKW ? 1 : 0;
      for (int f = threadIdx.x; f < nact * 128; f += blockDim.x) {     // 128 float4 per block
        const int q = f >> 7, p = (f >> 6) & 1, h = (f >> 5) & 1, l = f & 31;
        const float *r0 = L + (size_t)(4 * t[q] + 2 * p) * kNPix + c * kGvKW + 4 * l + 2 * h;
        const float2 x = *reinterpret_cast<const float2 *>(r0), y = *reinterpret_cast<const float2 *>(r0 + kNPix);
        dst[f] = make_float4(x.x, y.x, x.y, y.y);
      }
      dst += nact * 128;
    }
  }
}

cudaError_t launch_gemv_pack_L(const float *L, float *Lg, const int *table_dev, int n_ctas, int variant, int dense, cudaStream_t s) {
  (void)variant;
  gemv_pack_L_kernel<<<n_ctas, 256, 0, s>>>(L, Lg, table_dev, dense);
  return cudaGetLastError();
}

// the kernel-parameter form of a schedule (host memory, owned by the handle)
GvTable *gemv_make_table(const int *sched_host, int n_ctas, int dense) {
  if (n_ctas > kGvMaxCtas) return nullptr;
  GvTable *t = new (std::nothrow) GvTable();
  if (!t) return nullptr;
  memset(t, 0xff, sizeof(*t));
  t->max_blocks = 1;
  for (int c = 0; c < n_ctas; ++c) {
    t->off[c] = sched_host[c * kGemvTableStride + kGemvSlots];
    int kend[kGemvSlots], kmax = 0;
    for (int q = 0; q < kGemvSlots; ++q) {
      const int Q = sched_host[c * kGemvTableStride + q];
      t->quad[c][q] = (short)Q;
      kend[q] = Q < 0 ? 0 : (dense ? kNPix : 4 * Q + 4);
      kmax = std::max(kmax, kend[q]);
    }
    const int n_chunks = (kmax + kGvKW - 1) / kGvKW;
    for (int s = 0; 2 * s < n_chunks; ++s) {
      const int ca = s, cb = n_chunks - 1 - s;
      int nact = 0;
      for (int q = 0; q < kGemvSlots; ++q) nact += (kend[q] > ca * kGvKW ? 1 : 0) + ((cb > ca && kend[q] > cb * kGvKW) ? 1 : 0);
      t->max_blocks = std::max(t->max_blocks, nact);
    }
  }
  return t;
}
void gemv_free_table(GvTable *t) { delete t; }

static bool make_map_2d(CUtensorMap *m, const float *ptr, int rows, int box_k, int box_rows, CUtensorMapL2promotion promo) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)kNPix, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kNPix * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NC>
static cudaError_t launch_gv(const GemvArgs &g, cudaStream_t s) {
  using Cfg = GvCfg<NC>;
  GemvKernelArgs a;
  a.Lg = g.Lg;
  a.dense = g.dense;
  a.n_cols = g.n_cols;
  a.om = OutMap{g.z_cols, g.gamma, g.out, g.out_bn, g.out_wn, g.B, g.C, g.res_mode, g.train};
  a.trace = g.trace;
  const uint32_t kStatic = 1536 + (uint32_t)sizeof(OutPos4) * (Cfg::NOUT / 4);      // static shared memory of the kernel
  const uint32_t budget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - kStatic;
  if (!g.table || g.n_ctas > kGvMaxCtas) {
    set_error("gemv contraction: no schedule table (%d CTAs)", g.n_ctas);
    return cudaErrorInvalidValue;
  }
  a.zoff = (uint32_t)g.table->max_blocks * kGvQuadBytes;
  a.stage_bytes = a.zoff + 2 * Cfg::kZBytes;
  int stages = (int)(budget / a.stage_bytes);
  if (stages > 12) stages = 12;
  if (const char *e = getenv("BNDM_GV_STAGES")) {          // experiment knob
    const int x = atoi(e);
    if (x >= 2 && x <= stages) stages = x;
  }
  // the ring doubles as the partial-sum dump of the final reduction (+ the reduced outputs behind it)
  while ((size_t)stages * a.stage_bytes < (size_t)Cfg::kPartBytes + 16 + Cfg::kRedBytes) ++stages;
  a.stages = stages;
  const size_t smem = (size_t)stages * a.stage_bytes + 1024 + 2 * stages * 8 + 16;
  if (stages < 2 || smem > 227 * 1024 - kStatic) {
    set_error("gemv contraction: shared memory budget exceeded");
    return cudaErrorInvalidValue;
  }
  // (dynamic + static must stay within 227 KiB: ask for what is used, not for the maximum)
  cudaError_t e = cudaFuncSetAttribute(gemv_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  CUtensorMap map_z;
  memset(&map_z, 0, sizeof(map_z));
  if (!make_map_2d(&map_z, g.z_cols, g.n_cols, kGvKW, NC, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) {
    set_error("cuTensorMapEncodeTiled failed (gemv contraction, %d columns)", g.n_cols);
    return cudaErrorInvalidValue;
  }
  return launch_pdl(gemv_kernel<NC>, dim3(g.n_ctas), dim3(kGvThreads), smem, s, a, map_z, *g.table);
}

cudaError_t launch_gemv(const GemvArgs &g, cudaStream_t s) {
  if (g.n_cols < 1 || g.n_cols > kGemvMaxCols) {
    set_error("gemv contraction: %d columns (max %d)", g.n_cols, kGemvMaxCols);
    return cudaErrorInvalidValue;
  }
  if (g.n_cols <= 4) return launch_gv<4>(g, s);
  if (g.n_cols <= 8) return launch_gv<8>(g, s);
  if (g.n_cols <= 12) return launch_gv<12>(g, s);
  return launch_gv<16>(g, s);
}

}  // namespace bndm
