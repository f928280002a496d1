// Internal helpers shared by the kernels of libbndm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bndm {

constexpr int kTile = 64;            // blue-noise tile edge
constexpr int kNPix = kTile * kTile; // 4096 = M = K of the contraction
constexpr int kBlk = 128;            // row-tile / k-block edge of the triangular schedule
constexpr int kNumBlk = kNPix / kBlk;
constexpr int kStageK = 32;          // k extent of one tensor-core pipeline stage (128-byte rows)

// Output mapping of GEMM column j / row p (see DESIGN.md "K1 addressing")
enum ResMode : int { kRes64 = 0, kRes32 = 1, kRes128 = 2 };

void set_error(const char *fmt, ...);

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// The kernels of one get_noise call (pack -> contraction -> combine) are launched with
// programmatic stream serialisation: a kernel may start while its predecessor is still
// running (launch latency, barrier/TMEM set-up and the L loads of the contraction overlap the
// predecessor's tail) and calls pdl_wait() before it touches anything a predecessor wrote or
// still reads.  pdl_wait() returns once the preceding grid has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // BNDM_NO_PDL=1 turns it off (A/B measurements)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- the triangular unit schedule, shared by GEMM kernels and the epilogue ---------------
// Row tile i covers rows [128 i, 128 i + 128); it needs k-blocks [0, kb(i)) with
// kb(i) = i + 1 (lower-triangular L) or 32 (dense).  The k range is cut in chunks of `kc`
// k-blocks; every (row tile, chunk) is one work unit that writes one partial tile.
struct Schedule {
  int n_row_tiles;   // 32, or 16 for the 32^2 branch (rows with h >= 32 are never stored)
  int kc;            // k-blocks per unit
  int dense;         // 0: lower-triangular, 1: dense
  __host__ __device__ int kblocks(int i) const { return dense ? kNumBlk : i + 1; }
  __host__ __device__ int nsplit(int i) const { return (kblocks(i) + kc - 1) / kc; }
  __host__ __device__ int base(int i) const {
    int b = 0;
    for (int r = 0; r < i; ++r) b += nsplit(r);
    return b;
  }
  __host__ __device__ int n_units() const { return base(n_row_tiles); }
  // unit index -> (row tile, first k-block, number of k-blocks)
  __host__ __device__ void decode(int u, int &i, int &kb0, int &nkb) const {
    int r = 0;
    while (u >= nsplit(r)) { u -= nsplit(r); ++r; }
    i = r;
    kb0 = u * kc;
    int rem = kblocks(r) - kb0;
    nkb = rem < kc ? rem : kc;
  }
};

// ---- launchers (defined in the .cu files) -------------------------------------------------
struct PackArgs {
  const float *src;   // white source
  float *z_raw;       // [n_cols_pad][4096] packed raw columns (may be null if src is already packed)
  float *zt;          // tcgen05 operand: stage blocks [col block][128 stages][zh rows | zl rows][32 k],
                      // SWIZZLE_128B image (null for the SIMT path); see noise_gemm_tc.cu
  int nb;             // columns per column block of zt
  int n_cols;         // B*C*(tiles per image)
  int n_cols_pad;     // rows of the packed buffers (zero-filled above n_cols)
  int B, C;
  int res_mode;       // ResMode
  int src_is_image;   // BNDM_SRC_IMAGE
  int Bg = 0;         // 128^2 image source: GLOBAL batch of the image tensor (get_noise_recent.py:131-146 mixes samples
  int n0 = 0;         // across the batch: tile n = k*Bg + b); this call owns tiles [n0, n0 + 4 B)
};
cudaError_t launch_pack(const PackArgs &a, cudaStream_t s);

struct GemmArgs {
  const float *L;        // raw L (SIMT path)
  const float *z;        // packed raw z columns (SIMT path)
  float *partials;       // [n_units][n_cols_pad][128]
  int n_cols_pad;
  Schedule sched;
};
cudaError_t launch_gemm_simt(const GemmArgs &a, cudaStream_t s);

// Training-side fusion (iadb_bn.py:881-954): when x_alpha != null the epilogue also emits
//   x_alpha = alpha[b] * x0 + (1 - alpha[b]) * x1       (:915, x0 = the lerped noise, x1 = data)
//   tar1    = x1 - x0                                   (:949 / :976)
//   tar2    = alpha_prev[b] * (noise_bn - noise_wn)     (:950; null for 'GBN')
// in the same pass, with the reference's association (explicit _rn ops).
struct TrainOut {
  const float *x1, *alpha, *alpha_prev;
  float *x_alpha, *tar1, *tar2;
};

struct EpilogueArgs {
  const float *partials;
  const float *z_cols;   // packed raw columns [n_cols_pad][4096] (white values)
  const float *gamma;    // [B] or null
  float *out, *out_bn, *out_wn;
  int n_cols, n_cols_pad;
  int B, C;
  int res_mode;
  Schedule sched;
  TrainOut train;
};
cudaError_t launch_epilogue(const EpilogueArgs &a, cudaStream_t s);

cudaError_t launch_white128(const float *x, float *out, int B, int C, int Bg, int b0, cudaStream_t s);
// L -> stage blocks [Lh tile | Ll tile] in the SWIZZLE_128B shared-memory image (noise_gemm_tc.cu);
// n_blocks = 2112 (lower-triangular, blocks of row tile i start at 2 i (i + 1)) or 4096 (dense)
cudaError_t launch_tile_L(const float *L, float *Lt, int dense, int raw, cudaStream_t s);
inline size_t tile_L_blocks(int dense) { return dense ? (size_t)kNumBlk * (kNPix / kStageK) : (size_t)2 * kNumBlk * (kNumBlk + 1); }
constexpr size_t kLBlockFloats = 2 * kBlk * kStageK;   // 32 KiB per stage block

// byte offset of element (row r, k-in-stage kk) inside a K-major SWIZZLE_128B tile of 128-byte rows
__host__ __device__ inline uint32_t sw128_offset(int r, int kk) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((kk >> 2) ^ (r & 7)) & 7) << 4) + (kk & 3) * 4);
}
cudaError_t launch_tri_check(const float *L, int n, int *flag_dev, cudaStream_t s);

// ---- stream-K schedule of the tensor-core contraction (see noise_gemm_tc.cu) ---------------
// Stage = (row tile, 32 k values, column block).  Row tile i owns 4 (i + 1) stages (lower-
// triangular L) or 128 (dense); the W stages of all column blocks are cut into G contiguous
// equal ranges, one per CTA.  Inside a column block the row tiles are visited from the LAST
// (longest) to the first, so the work that finishes last belongs to the shortest tiles and the
// fused combine of the big tiles overlaps the rest of the kernel.  Shared by the GEMM kernel
// and the combine kernel.
struct StreamK {
  int n_tiles;    // row tiles: 32, or 16 for the 32^2 branch (rows with h >= 32 are never stored)
  int dense;      // 0: lower-triangular, 1: dense
  int n_colblk;   // column blocks of nb columns
  int G;          // CTAs = min(#SMs, W)
  int Stot;       // units per column block
  int W;          // n_colblk * Stot
  int sub;        // k-stages (32 k each) per schedule unit / pipeline stage: 1 or 2
  // units of row tiles [0, i); times `sub` = index of tile i's first stage block in Lt
  __host__ __device__ int cum(int i) const { return (dense ? (kNPix / kStageK) * i : 2 * i * (i + 1)) / sub; }
  __host__ __device__ int cta_begin(int c) const { return (int)((int64_t)c * W / G); }
  __host__ __device__ int cta_of(int g) const { return (int)((((int64_t)g + 1) * G - 1) / W); }
  __host__ __device__ int tile_begin(int cb, int tile) const { return cb * Stot + Stot - cum(tile + 1); }
  __host__ __device__ int tile_end(int cb, int tile) const { return cb * Stot + Stot - cum(tile); }
  __host__ __device__ void decode(int g, int &cb, int &tile, int &s) const {
    cb = g / Stot;
    const int r = g - cb * Stot;
    const int a = Stot - 1 - r;          // position in ascending-tile order
    int i = 0;
    while (cum(i + 1) <= a) ++i;
    tile = i;
    s = r - (Stot - cum(i + 1));         // stage inside the tile, ascending k
  }
  // partial-tile slot of the segment CTA `cta` computes inside (cb, tile): cta + visit index is
  // strictly increasing along the global stage order, hence unique per segment
  __host__ __device__ int slot(int cta, int cb, int tile) const { return cta + cb * n_tiles + (n_tiles - 1 - tile); }
  __host__ __device__ int n_slots() const { return G + n_colblk * n_tiles - 1; }
  // CTAs that contribute a segment to (cb, tile): first .. last, ascending k
  __host__ __device__ int first_cta(int cb, int tile) const { return cta_of(tile_begin(cb, tile)); }
  __host__ __device__ int last_cta(int cb, int tile) const { return cta_of(tile_end(cb, tile) - 1); }
};
inline StreamK make_streamk(int n_tiles, int dense, int n_colblk, int num_sms, int sub = 1) {
  StreamK k;
  k.sub = sub;
  k.n_tiles = n_tiles;
  k.dense = dense;
  k.n_colblk = n_colblk;
  k.Stot = k.cum(n_tiles);
  k.W = n_colblk * k.Stot;
  k.G = num_sms < k.W ? num_sms : k.W;
  return k;
}

struct TcGemmArgs {
  const float *Lt;            // stage blocks of L (launch_tile_L), triangular or dense per sk.dense
  int raw_L;                  // Lt holds raw fp32 16 KiB blocks (converter variant) instead of hi|lo 32 KiB blocks
  const float *zt;            // stage blocks of z (launch_pack)
  float *partials;            // [n_slots][nb][128]
  int n_cols_pad;             // n_colblk * nb
  int nb;                     // columns per column block (multiple of 16, <= 128)
  unsigned long long *trace;  // debug time stamps [cta][24] or null
  StreamK sk;
  // fused combine (the last CTA to finish a row tile sums its partial tiles and writes the
  // outputs; no combine kernel): out == null -> partial tiles only
  int *tile_counters;         // [n_colblk][n_tiles], zero between calls (self-resetting)
  const float *z_cols;        // packed raw columns [.][4096] (white values)
  const float *gamma;         // [B] or null
  float *out, *out_bn, *out_wn;
  int n_cols, B, C, res_mode;
};
cudaError_t launch_gemm_tc(const TcGemmArgs &a, cudaStream_t s);
int tc_pick_nb(int n_cols);   // column block for a given column count
int tc_num_sms();
bool tc_fused_combine(int nb);
bool tc_raw_L(int nb, int n_colblk);   // policy: raw-L converter variant for this shape?
int tc_sub(int nb, bool raw);          // policy: k-stages per pipeline stage (2 = 32 KiB L requests)
void tc_set_policy(int fused, int raw);   // policy: combine fused into the contraction for this column block?

// ---- K1g: streaming fp32 contraction for <= kGemvMaxCols columns (noise_gemv.cu) -------------
constexpr int kGemvSlots = 8;      // quads (4 consecutive rows of L) a CTA may own
constexpr int kGemvMaxCols = 16;
constexpr int kGemvAutoCols = 16;  // the default rule picks K1g up to here (measured: 18.4 us at 16 columns, K1b 19.8)
constexpr int kGemvTraceStride = 128;   // u64 per CTA of K1g's debug trace: [0..3] summary, [8+c] stage c requested, [48+c] landed, [88+c] released
constexpr int kGemvTableStride = 10;   // ints per CTA in the schedule table: 8 slots, first stream block, load
struct GvTable;            // a schedule in its kernel-parameter form (noise_gemv.cu)
GvTable *gemv_make_table(const int *sched_host, int n_ctas, int dense);     // null if the device has too many SMs for the parameter block
void gemv_free_table(GvTable *t);
struct GemvArgs {
  const float *Lg;         // L in K1g's stream order (launch_gemv_pack_L)
  const float *z_cols;     // white columns [n_cols][4096]
  const GvTable *table;    // host, the schedule of this row set (gemv_make_table)
  int n_ctas;
  int dense;               // L is not lower-triangular (or the caller forces the dense walk)
  int variant;             // gemv_variant()
  const float *gamma;      // [B] or null
  float *out, *out_bn, *out_wn;
  int n_cols, B, C, res_mode;
  TrainOut train;
  unsigned long long *trace;
};
cudaError_t launch_gemv(const GemvArgs &a, cudaStream_t s);
long gemv_build_schedule(int res32, int dense, int n_ctas, int variant, int *table);   // -> blocks of 4 x KW floats, -1: no fit
cudaError_t launch_gemv_pack_L(const float *L, float *Lg, const int *table_dev, int n_ctas, int variant, int dense, cudaStream_t s);
int gemv_kw(int variant);
int gemv_variant();
bool gemv_policy();                 // BNDM_GEMV=0 turns the automatic choice of K1g off (A/B measurements)

// combine of the stream-K partials + everything get_noise_v2 does after the matmul
struct CombineArgs {
  const float *partials;   // [n_slots][nb][128]
  const float *z_cols;     // packed raw columns [.][4096] (white values)
  const float *gamma;      // [B] or null
  float *out, *out_bn, *out_wn;
  int n_cols, nb;
  int B, C;
  int res_mode;
  StreamK sk;
  TrainOut train;
};
cudaError_t launch_combine(const CombineArgs &a, cudaStream_t s);

struct IadbArgs {
  float *x_out;
  const float *x, *d;
  const float *dalpha, *dgamma;   // per-sample [B]                       (direct mode)
  const float *table;             // rows {dalpha, dgamma, t_next, 0}     (scheduled mode)
  int *state;                     // {step_idx, blocks_done}              (scheduled mode)
  float *t_next_out;              // [B] or null
  int B, C, HW, Cd;
  int d_nhwc;                     // d is channels-last [B][HW][Cd] (the fused UNet's native output) instead of NCHW
};

cudaError_t launch_iadb_step(const IadbArgs &a, bool sched, cudaStream_t s);

struct DdimArgs {
  float *x_out;
  const float *x, *eps, *noise;
  const float *coef;   // rows of 8
  int *state;          // may be null
  float *t_next_out;
  int B, clip;
  int64_t n;
};

cudaError_t launch_ddim_step(const DdimArgs &a, cudaStream_t s);
cudaError_t launch_upsample2x_nhwc(const float *x, float *y, int B, int H, int W, int C, cudaStream_t s);
cudaError_t launch_attention_small(const float *qkv, float *out, int B, int T, int C, int head_dim, cudaStream_t s);
cudaError_t launch_add_bias_nhwc(const float *a, const float *a2, const float *bias_a, const float *b, const float *bias_b,
                                 float *out, size_t n, int C, cudaStream_t s);
cudaError_t launch_groupnorm_nhwc(const float *x, const float *x2, int C1, const float *res, const float *add_bc, int add_stride,
                                  const float *weight,
                                  const float *bias, float *sum_out, float *y, int B, int C, int HW, int groups, float eps,
                                  int silu, cudaStream_t s);
// K9 (linear_tc.cu): out[m][n] = sum_k a[m][k] w[n][k] (+ bias[n]) as a 3xTF32 tcgen05 GEMM
cudaError_t launch_linear_tc(const float *a, const float *w, const float *bias, float *out, int M, int N, int K, cudaStream_t s);
// K10 (shortcut_tc.cu): out[m][n] = sum_k w[n][k] cat(x, x2)[m][k] + h2[m][n] + bias[n] (1x1 convolution in TF32 + residual + biases)
cudaError_t launch_shortcut_tc(const float *x, const float *x2, int C1, int C2, const float *w, const float *h2, const float *bias,
                               float *out, long long M, int N, cudaStream_t s);
// K11 (conv_in.cu): 3x3 convolution of an NCHW input with <= 4 channels to an NHWC activation (no bias)
cudaError_t launch_conv_in3x3(const float *x, const float *w, float *out, int B, int Cin, int H, int W, int Cout, cudaStream_t s);
cudaError_t launch_to_u8(const float *x, uint8_t *out, int B, int C, int HW, cudaStream_t s);
cudaError_t launch_snapshot_u8(const float *x, uint8_t *out, int N, int C, int HW, const int *final_flags, int final_all,
                               cudaStream_t s);

// ---- tf32 split (round-to-nearest-away hi, residual lo) -----------------------------------
__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float v, float &hi, float &lo) {
  hi = tf32_rna(v);
  lo = tf32_rna(__fsub_rn(v, hi));
}

}  // namespace bndm
