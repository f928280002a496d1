// C ABI of libbndm_b200.so -- see include/bndm_b200.h for the contract of every entry point.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <new>

#include "../../include/bndm_b200.h"
#include "common.cuh"

namespace bndm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int fail_cuda(cudaError_t e, const char *what) {
  char detail[256];
  snprintf(detail, sizeof(detail), "%s", g_err);          // a launcher may have left the precise reason
  set_error("%s: %s%s%s", what, cudaGetErrorString(e), detail[0] ? " -- " : "", detail);
  return BNDM_ERR_CUDA;
}

#define CK(expr)                                        \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
  } while (0)

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("BNDM_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

static bool stream_is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st != cudaStreamCaptureStatusNone;
}

}  // namespace bndm

using namespace bndm;

struct bndm_L {
  const float *L = nullptr;     // caller's matrix (bound, not owned)
  float *Lt = nullptr;          // owned: tcgen05 operand, tf32 hi|lo stage blocks (triangular set if L is lower-triangular)
  float *Lr = nullptr;          // owned: raw fp32 stage blocks of a triangular L (converter variant, 16 KiB each)
  float *Lt_dense = nullptr;    // owned: all 4096 stage blocks of a triangular L, built on the first BNDM_FORCE_DENSE call
  int n = 0;
  int lower_triangular = 0;
  int sm100 = 0;
  // workspace (owned), sized for `cap_cols` padded columns
  int cap_cols = 0;
  int req_cols = 0;             // max_columns the workspace was sized for
  size_t cap_partial = 0;
  float *z_raw = nullptr, *zt = nullptr, *partials = nullptr;
  int *tile_counters = nullptr;   // owned: kMaxTileCounters ints, zero between calls (fused combine)
  int64_t ws_bytes = 0;
  // optional per-launch timing (bndm_profile_enable)
  int *gv_sched[4] = {nullptr, nullptr, nullptr, nullptr};   // owned: K1g row schedules [2 res32 + dense] -> [n_sms][kGemvTableStride]
  float *gv_L[4] = {nullptr, nullptr, nullptr, nullptr};     // owned: L in the stream order of that schedule
  GvTable *gv_tab[4] = {nullptr, nullptr, nullptr, nullptr}; // owned (host): the schedule as kernel parameters
  int gv_variant = 0;
  int n_sms = 148;
  unsigned long long *trace = nullptr;   // debug: per-CTA time stamps of the contraction kernel (caller-owned)
  int profile = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int ev_valid = 0;
  // one workspace per handle: calls are serialised (host threads: mutex; streams: event hand-over)
  std::mutex mu;
  cudaStream_t last_stream = nullptr;
  int has_last = 0;
  cudaEvent_t xstream = nullptr;
};

static const int kMaxTileCounters = 4096;
static const int kUnitCap = 592;   // most partial tiles a SIMT schedule may produce per column block

static int pad_cols(int n_cols) {
  const int nb = tc_pick_nb(n_cols);
  return (n_cols + nb - 1) / nb * nb;
}

// SIMT reference path: chunk length of the split-K schedule (enough units to fill the SMs for
// the given number of column blocks, never more than kUnitCap partial tiles).
static Schedule pick_schedule(int n_row_tiles, int dense, int col_blocks) {
  Schedule s;
  s.n_row_tiles = n_row_tiles;
  s.dense = dense;
  for (int kc = 8; kc >= 1; --kc) {
    s.kc = kc;
    if (s.n_units() * col_blocks >= 148) break;
  }
  while (s.n_units() > kUnitCap) ++s.kc;
  return s;
}

// fp32 elements of partial-tile workspace a call with n_cols columns may need (any branch,
// either contraction kernel)
static size_t partial_elems(int n_cols) {
  const int cp = pad_cols(n_cols);
  const int nb = tc_pick_nb(n_cols);
  size_t worst = 0;
  for (int dense = 0; dense < 2; ++dense)
    for (int rows = kNumBlk / 2; rows <= kNumBlk; rows += kNumBlk / 2) {
      const Schedule s = pick_schedule(rows, dense, (cp + 63) / 64);
      const size_t e_simt = (size_t)s.n_units() * cp * kBlk;
      const StreamK k = make_streamk(rows, dense, cp / nb, tc_num_sms());
      const size_t e_tc = (size_t)k.n_slots() * nb * kBlk;
      if (e_simt > worst) worst = e_simt;
      if (e_tc > worst) worst = e_tc;
    }
  return worst;
}

static void free_ws(bndm_L *h) {
  cudaFree(h->z_raw);
  cudaFree(h->zt);
  cudaFree(h->partials);
  h->z_raw = h->zt = h->partials = nullptr;
  h->cap_cols = 0;
  h->req_cols = 0;
  h->cap_partial = 0;
  h->ws_bytes = 0;
}

static int alloc_ws(bndm_L *h, int max_columns) {
  free_ws(h);
  const int cols_pad = pad_cols(max_columns);
  size_t pe = 0;
  // a smaller call may pick a finer schedule / another column blocking: size for the worst of
  // every column count up to the requested one (cheap: a few hundred evaluations)
  for (int n = 1; n <= max_columns; n = (n < 16 ? n + 1 : n + 16)) {
    const size_t e = partial_elems(n);
    if (e > pe) pe = e;
  }
  {
    const size_t e = partial_elems(max_columns);
    if (e > pe) pe = e;
  }
  const size_t zb = (size_t)cols_pad * kNPix * sizeof(float);
  CK(cudaMalloc(&h->z_raw, zb));
  CK(cudaMalloc(&h->zt, 2 * zb));
  CK(cudaMalloc(&h->partials, pe * sizeof(float)));
  h->cap_cols = cols_pad;
  h->req_cols = max_columns;
  h->cap_partial = pe;
  h->ws_bytes = (int64_t)(3 * zb + pe * sizeof(float));
  return BNDM_OK;
}

static int reserve_unlocked(bndm_L *h, int max_columns, void *stream);

// K1g set-up for table k = 2 res32 + dense: host row schedule -> device, then L gathered into stream order.
// Leaves gv_sched[k] null (K1g unavailable, K1b takes over) when the quads do not fit the device's SMs.
static int gv_build(bndm_L *h, int k, cudaStream_t s) {
  if (h->gv_sched[k]) return BNDM_OK;
  const size_t n = (size_t)h->n_sms * kGemvTableStride;
  int *host = new (std::nothrow) int[n];
  if (!host) { set_error("out of host memory"); return BNDM_ERR_ARG; }
  const long blocks = gemv_build_schedule(k >> 1, k & 1, h->n_sms, h->gv_variant, host);
  if (blocks < 0) { delete[] host; return BNDM_OK; }
  int *sched = nullptr;
  float *Lg = nullptr;
  cudaError_t e = cudaMalloc(&sched, n * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&Lg, (size_t)blocks * 4 * gemv_kw(h->gv_variant) * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpyAsync(sched, host, n * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = launch_gemv_pack_L(h->L, Lg, sched, h->n_sms, h->gv_variant, k & 1, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  GvTable *tab = e == cudaSuccess ? gemv_make_table(host, h->n_sms, k & 1) : nullptr;
  delete[] host;
  if (e != cudaSuccess) { cudaFree(sched); cudaFree(Lg); return fail_cuda(e, "K1g schedule / stream-ordered L"); }
  if (!tab) { cudaFree(sched); cudaFree(Lg); return BNDM_OK; }     // more SMs than the parameter block holds: K1b takes over
  h->gv_sched[k] = sched;
  h->gv_L[k] = Lg;
  h->gv_tab[k] = tab;
  return BNDM_OK;
}

extern "C" {

int bndm_version(void) { return BNDM_ABI_VERSION; }
const char *bndm_last_error(void) { return g_err; }

int bndm_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int bndm_prepare_L(const float *L_dev, int n, int max_columns, void *stream, bndm_L **out) {
  if (!L_dev || !out) { set_error("bndm_prepare_L: null argument"); return BNDM_ERR_ARG; }
  if (n != kNPix) { set_error("bndm_prepare_L: cov_mat_L must be (4096,4096), got n=%d", n); return BNDM_ERR_UNSUPPORTED; }
  if (max_columns < 1) max_columns = 1;
  cudaStream_t s = (cudaStream_t)stream;
  bndm_L *h = new (std::nothrow) bndm_L();
  if (!h) { set_error("out of host memory"); return BNDM_ERR_ARG; }
  h->L = L_dev;
  h->n = n;
  h->sm100 = bndm_device_is_sm100();

  int *flag = nullptr;
  cudaError_t e = cudaMalloc(&flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, sizeof(int), s);
  if (e == cudaSuccess) e = launch_tri_check(L_dev, n, flag, s);
  int host_flag = 1;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(flag);
  if (e != cudaSuccess) { delete h; return fail_cuda(e, "triangularity check"); }
  h->lower_triangular = host_flag ? 0 : 1;

  if (h->sm100) {
    const int dense = h->lower_triangular ? 0 : 1;
    e = cudaMalloc(&h->Lt, tile_L_blocks(dense) * kLBlockFloats * sizeof(float));
    if (e == cudaSuccess) e = launch_tile_L(L_dev, h->Lt, dense, 0, s);
    if (e == cudaSuccess && !dense) {
      e = cudaMalloc(&h->Lr, tile_L_blocks(0) * (kLBlockFloats / 2) * sizeof(float));
      if (e == cudaSuccess) e = launch_tile_L(L_dev, h->Lr, 0, 1, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { bndm_free_L(h); return fail_cuda(e, "tf32 split / tiling of L"); }
  }
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || h->n_sms < 1)
      h->n_sms = 148;
    h->gv_variant = gemv_variant();
    // K1g: row schedule + stream-ordered copy of L for the 64^2/128^2 and the 32^2 row sets (the dense walk of a
    // triangular L is a testing path, built on first use)
    if (h->sm100)
      for (int res32 = 0; res32 < 2; ++res32) {
        int rc = gv_build(h, 2 * res32 + (h->lower_triangular ? 0 : 1), s);
        if (rc != BNDM_OK) { bndm_free_L(h); return rc; }
      }
  }
  int rc = alloc_ws(h, max_columns);
  if (rc != BNDM_OK) { bndm_free_L(h); return rc; }
  e = cudaMalloc(&h->tile_counters, kMaxTileCounters * sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(h->tile_counters, 0, kMaxTileCounters * sizeof(int), s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { bndm_free_L(h); return fail_cuda(e, "tile counters"); }
  *out = h;
  return BNDM_OK;
}

int bndm_reserve_columns(bndm_L *h, int max_columns, void *stream) {
  if (!h) { set_error("null handle"); return BNDM_ERR_ARG; }
  std::lock_guard<std::mutex> lock(h->mu);
  return reserve_unlocked(h, max_columns, stream);
}

}  // extern "C"

static int reserve_unlocked(bndm_L *h, int max_columns, void *stream) {
  if (max_columns <= h->req_cols) return BNDM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (stream_is_capturing(s)) { set_error("workspace growth requested during stream capture"); return BNDM_ERR_WORKSPACE; }
  CK(cudaStreamSynchronize(s));
  return alloc_ws(h, max_columns);
}

extern "C" {

int bndm_L_is_lower_triangular(const bndm_L *h) { return h ? h->lower_triangular : 0; }
int64_t bndm_workspace_bytes(const bndm_L *h) { return h ? h->ws_bytes : 0; }

int bndm_profile_enable(bndm_L *h, int on) {
  if (!h) { set_error("null handle"); return BNDM_ERR_ARG; }
  if (on && !h->ev[0])
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&h->ev[i]));
  h->profile = on ? 1 : 0;
  h->ev_valid = 0;
  return BNDM_OK;
}

int bndm_profile_last_ms(bndm_L *h, float *pack_ms, float *gemm_ms, float *epilogue_ms) {
  if (!h || !h->ev_valid) { set_error("no profiled call recorded"); return BNDM_ERR_ARG; }
  CK(cudaEventSynchronize(h->ev[3]));
  float t[3];
  for (int i = 0; i < 3; ++i) CK(cudaEventElapsedTime(&t[i], h->ev[i], h->ev[i + 1]));
  if (pack_ms) *pack_ms = t[0];
  if (gemm_ms) *gemm_ms = t[1];
  if (epilogue_ms) *epilogue_ms = t[2];
  return BNDM_OK;
}

int bndm_debug_set_policy(int fused_combine, int raw_L) {
  tc_set_policy(fused_combine, raw_L);
  return BNDM_OK;
}

int bndm_debug_set_trace(bndm_L *h, unsigned long long *trace_dev) {
  if (!h) { set_error("null handle"); return BNDM_ERR_ARG; }
  h->trace = trace_dev;
  return BNDM_OK;
}

int bndm_free_L(bndm_L *h) {
  if (!h) return BNDM_OK;
  for (int i = 0; i < 4; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->xstream) cudaEventDestroy(h->xstream);
  free_ws(h);
  cudaFree(h->Lt);
  cudaFree(h->Lt_dense);
  cudaFree(h->Lr);
  cudaFree(h->tile_counters);
  for (int k = 0; k < 4; ++k) { cudaFree(h->gv_sched[k]); cudaFree(h->gv_L[k]); gemv_free_table(h->gv_tab[k]); }
  delete h;
  return BNDM_OK;
}

}  // extern "C"

static int get_noise_impl(bndm_L *h, const float *z, const float *gamma, float *out, float *out_bn, float *out_wn, int B,
                          int C, int res, unsigned flags, void *stream, const TrainOut &train, int Bg = 0, int b0 = 0) {
  g_err[0] = 0;
  if (!h || !z || (!out && !train.x_alpha)) { set_error("bndm_get_noise_f32: null argument"); return BNDM_ERR_ARG; }
  if (B < 1 || C < 1) { set_error("bndm_get_noise_f32: bad shape B=%d C=%d", B, C); return BNDM_ERR_ARG; }
  if (Bg == 0) Bg = B;                       // not sharded
  if (b0 < 0 || b0 + B > Bg) { set_error("bndm_get_noise_shard_f32: shard [%d, %d) outside the global batch %d", b0, b0 + B, Bg); return BNDM_ERR_ARG; }
  int mode;
  if (res == 64) mode = kRes64;
  else if (res == 32) mode = kRes32;
  else if (res == 128) mode = kRes128;
  else { set_error("bndm_get_noise_f32: resolution %d not implemented (32/64/128)", res); return BNDM_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->has_last && h->last_stream != s) {
    // the previous call used another stream and may still be running on the shared workspace: this stream waits
    // for everything submitted there so far
    if (stream_is_capturing(h->last_stream)) {
      set_error("bndm_L handle used on a second stream while its previous stream is being captured (one handle owns one workspace)");
      return BNDM_ERR_WORKSPACE;
    }
    if (!stream_is_capturing(s)) {
      if (!h->xstream) CK(cudaEventCreateWithFlags(&h->xstream, cudaEventDisableTiming));
      CK(cudaEventRecord(h->xstream, h->last_stream));
      CK(cudaStreamWaitEvent(s, h->xstream, 0));
    }
    // (a capture starting on another stream cannot wait for work outside it: the capture protocol already requires
    // the caller to have ordered earlier work before the capturing stream, e.g. side.wait_stream(current))
  }
  h->last_stream = s;
  h->has_last = 1;
  const int src_is_image = (flags & BNDM_SRC_IMAGE) ? 1 : 0;
  // Sharded call: z is the GLOBAL white field.  Only the 128^2 image source couples samples across the batch
  // (tile n = k*Bg + b is re-read as (b', k') = divmod(n, 4), get_noise_recent.py:131-146): there the gather kernel
  // works on global tile indices; everywhere else a shard's columns are a contiguous slice of the global ones.
  if (!(mode == kRes128 && src_is_image))
    z += (int64_t)b0 * C * (mode == kRes128 ? 4 * kNPix : (src_is_image && mode == kRes32) ? 1024 : kNPix);
  const bool simt = (flags & BNDM_GEMM_SIMT) != 0;
  if (!simt && !h->sm100) { set_error("the tcgen05 / TMA paths need an sm_100 device"); return BNDM_ERR_ARCH; }
  const int dense = (!h->lower_triangular || (flags & BNDM_FORCE_DENSE)) ? 1 : 0;

  const int n_cols = B * C * (mode == kRes128 ? 4 : 1);
  // K1g (streaming fp32 kernel, one launch, no split-K) in the GEMV regime; K1b (tcgen05) above it
  const int gv_table = (mode == kRes32 ? 2 : 0) + dense;
  bool gemv = false;
  if (dense && h->lower_triangular && h->sm100 && !h->gv_sched[gv_table] && !simt && !(flags & BNDM_GEMM_TC) && n_cols <= kGemvMaxCols &&
      !stream_is_capturing(s)) {
    int rc = gv_build(h, gv_table, s);             // BNDM_FORCE_DENSE on a triangular L (testing): built once
    if (rc != BNDM_OK) return rc;
  }
  if (flags & BNDM_GEMM_GEMV) {
    if (simt || (flags & BNDM_GEMM_TC)) { set_error("bndm_get_noise_f32: conflicting kernel flags"); return BNDM_ERR_ARG; }
    if (n_cols > kGemvMaxCols || !h->gv_sched[gv_table]) {
      set_error("BNDM_GEMM_GEMV: %d columns (max %d)", n_cols, kGemvMaxCols);
      return BNDM_ERR_UNSUPPORTED;
    }
    gemv = true;
  } else if (!simt && !(flags & BNDM_GEMM_TC)) {
    gemv = n_cols <= kGemvAutoCols && h->gv_sched[gv_table] != nullptr && gemv_policy();
  }
  if (gemv) {
    // white columns: the caller's tensor unless they must be gathered from an image (32^2 tiling, 128^2 quadrants)
    const bool gather = src_is_image && mode != kRes64;
    const int nbg = tc_pick_nb(n_cols);
    const int n_cols_pad_g = (n_cols + nbg - 1) / nbg * nbg;
    if (gather && n_cols_pad_g > h->cap_cols) {
      int rc = reserve_unlocked(h, n_cols, stream);
      if (rc != BNDM_OK) return rc;
    }
    if ((reinterpret_cast<uintptr_t>(z) % 16) != 0) { set_error("bndm_get_noise_f32: z must be 16-byte aligned"); return BNDM_ERR_ARG; }
    const bool prof = h->profile && !stream_is_capturing(s);
    if (prof) CK(cudaEventRecord(h->ev[0], s));
    if (gather) {
      PackArgs p;
      p.src = z;
      p.z_raw = h->z_raw;
      p.zt = nullptr;
      p.nb = nbg;
      p.n_cols = n_cols;
      p.n_cols_pad = n_cols_pad_g;
      p.B = B;
      p.C = C;
      p.res_mode = mode;
      p.src_is_image = src_is_image;
      p.Bg = Bg;
      p.n0 = 4 * b0;
      CK(launch_pack(p, s));
    }
    if (prof) CK(cudaEventRecord(h->ev[1], s));
    GemvArgs g;
    g.Lg = h->gv_L[gv_table];
    g.z_cols = gather ? h->z_raw : z;
    g.table = h->gv_tab[gv_table];
    g.n_ctas = h->n_sms;
    g.dense = dense;
    g.variant = h->gv_variant;
    g.gamma = gamma;
    g.out = out;
    g.out_bn = out_bn;
    g.out_wn = out_wn;
    g.n_cols = n_cols;
    g.B = B;
    g.C = C;
    g.res_mode = mode;
    g.train = train;
    g.trace = h->trace;
    CK(launch_gemv(g, s));
    if (prof) {
      CK(cudaEventRecord(h->ev[2], s));
      CK(cudaEventRecord(h->ev[3], s));
      h->ev_valid = 1;
    }
    return BNDM_OK;
  }
  const int nb = tc_pick_nb(n_cols);
  const int n_cols_pad = (n_cols + nb - 1) / nb * nb;
  const int n_row_tiles = mode == kRes32 ? kNumBlk / 2 : kNumBlk;
  const Schedule sched = pick_schedule(n_row_tiles, dense, (n_cols_pad + 63) / 64);          // SIMT path
  // raw-operand variant of the contraction: takes fp32 columns as they are (caller's tensor or the
  // gathered z_raw) and converts in shared memory -> the pack kernel only runs to gather
  const bool raw = !simt && !dense && h->Lr && tc_raw_L(nb, n_cols_pad / nb) && (reinterpret_cast<uintptr_t>(z) % 16 == 0);
  const StreamK sk = make_streamk(n_row_tiles, dense, n_cols_pad / nb, tc_num_sms(), tc_sub(nb, raw));   // tcgen05 path
  const size_t need = simt ? (size_t)sched.n_units() * n_cols_pad * kBlk : (size_t)sk.n_slots() * nb * kBlk;
  if (n_cols_pad > h->cap_cols || need > h->cap_partial) {
    int rc = reserve_unlocked(h, n_cols > h->req_cols ? n_cols : h->req_cols + 1, stream);
    if (rc != BNDM_OK) return rc;
    if (n_cols_pad > h->cap_cols || need > h->cap_partial) {
      set_error("internal: workspace still too small after growth (cols %d, partial %zu)", n_cols_pad, need);
      return BNDM_ERR_WORKSPACE;
    }
  }

  // K1a: the white columns are already in GEMM order unless they are gathered from an image;
  // the SIMT kernel reads zero-padded columns, so it always goes through the workspace copy.
  const bool src_packed = !(src_is_image && mode != kRes64);
  const bool need_raw = simt || !src_packed;
  PackArgs p;
  p.src = z;
  p.z_raw = need_raw ? h->z_raw : nullptr;
  p.zt = (simt || raw) ? nullptr : h->zt;
  p.nb = nb;
  p.n_cols = n_cols;
  p.n_cols_pad = n_cols_pad;
  p.B = B;
  p.C = C;
  p.res_mode = mode;
  p.src_is_image = src_is_image;
  p.Bg = Bg;
  p.n0 = 4 * b0;
  const bool prof = h->profile && !stream_is_capturing(s);
  if (prof) CK(cudaEventRecord(h->ev[0], s));
  if (p.z_raw || p.zt) CK(launch_pack(p, s));
  if (prof) CK(cudaEventRecord(h->ev[1], s));
  const float *z_cols = need_raw ? h->z_raw : z;

  if (simt) {
    // K1b (fp32 FFMA reference kernel): split-K partial tiles, then ordered combine + lerp + layout
    GemmArgs g;
    g.L = h->L;
    g.z = z_cols;
    g.partials = h->partials;
    g.n_cols_pad = n_cols_pad;
    g.sched = sched;
    CK(launch_gemm_simt(g, s));
    if (prof) CK(cudaEventRecord(h->ev[2], s));
    EpilogueArgs ep;
    ep.partials = h->partials;
    ep.z_cols = z_cols;
    ep.gamma = gamma;
    ep.out = out;
    ep.out_bn = out_bn;
    ep.out_wn = out_wn;
    ep.n_cols = n_cols;
    ep.n_cols_pad = n_cols_pad;
    ep.B = B;
    ep.C = C;
    ep.res_mode = mode;
    ep.sched = sched;
    ep.train = train;
    CK(launch_epilogue(ep, s));
  } else {
    // K1b (tcgen05, persistent stream-K) -> K1c (ordered combine + lerp + layout)
    TcGemmArgs g;
    g.Lt = h->Lt;
    g.raw_L = 0;
    if (raw) {
      g.Lt = h->Lr;
      g.raw_L = 1;
    }
    if (dense && h->lower_triangular) {
      // BNDM_FORCE_DENSE on a triangular L (testing): needs the full block set, built once
      if (!h->Lt_dense) {
        if (stream_is_capturing(s)) { set_error("dense operand copy requested during stream capture"); return BNDM_ERR_WORKSPACE; }
        CK(cudaMalloc(&h->Lt_dense, tile_L_blocks(1) * kLBlockFloats * sizeof(float)));
        CK(launch_tile_L(h->L, h->Lt_dense, 1, 0, s));
      }
      g.Lt = h->Lt_dense;
    }
    g.zt = h->zt;
    g.trace = h->trace;
    // small column blocks: the combine is fused into the contraction (the last CTA to finish a
    // row tile sums its partial tiles); large ones keep the wide combine kernel
    const bool fused = !train.x_alpha && tc_fused_combine(nb) && sk.n_colblk * sk.n_tiles <= kMaxTileCounters;
    g.tile_counters = h->tile_counters;
    g.z_cols = z_cols;
    g.gamma = gamma;
    g.out = fused ? out : nullptr;
    g.out_bn = out_bn;
    g.out_wn = out_wn;
    g.n_cols = n_cols;
    g.B = B;
    g.C = C;
    g.res_mode = mode;
    g.partials = h->partials;
    g.n_cols_pad = n_cols_pad;
    g.nb = nb;
    g.sk = sk;
    CK(launch_gemm_tc(g, s));
    if (prof) CK(cudaEventRecord(h->ev[2], s));
    if (fused) {
      if (prof) {
        CK(cudaEventRecord(h->ev[3], s));
        h->ev_valid = 1;
      }
      return BNDM_OK;
    }
    CombineArgs cb;
    cb.partials = h->partials;
    cb.z_cols = z_cols;
    cb.gamma = gamma;
    cb.out = out;
    cb.out_bn = out_bn;
    cb.out_wn = out_wn;
    cb.n_cols = n_cols;
    cb.nb = nb;
    cb.B = B;
    cb.C = C;
    cb.res_mode = mode;
    cb.sk = sk;
    cb.train = train;
    CK(launch_combine(cb, s));
  }
  if (prof) {
    CK(cudaEventRecord(h->ev[3], s));
    h->ev_valid = 1;
  }
  return BNDM_OK;
}

extern "C" {

int bndm_get_noise_f32(bndm_L *h, const float *z, const float *gamma, float *out, float *out_bn, float *out_wn, int B,
                       int C, int res, unsigned flags, void *stream) {
  if (!out) { set_error("bndm_get_noise_f32: null argument"); return BNDM_ERR_ARG; }
  return get_noise_impl(h, z, gamma, out, out_bn, out_wn, B, C, res, flags, stream,
                        TrainOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr});
}

int bndm_get_noise_train_f32(bndm_L *h, const float *z, const float *gamma, const float *x1, const float *alpha,
                             const float *alpha_prev, float *x_alpha, float *tar1, float *tar2, float *x0, int B, int C, int res,
                             unsigned flags, void *stream) {
  if (!x1 || !alpha || !x_alpha || !tar1) { set_error("bndm_get_noise_train_f32: null argument"); return BNDM_ERR_ARG; }
  if (tar2 && (!alpha_prev || !gamma)) { set_error("bndm_get_noise_train_f32: tar2 needs alpha_prev and gamma"); return BNDM_ERR_ARG; }
  return get_noise_impl(h, z, gamma, x0, nullptr, nullptr, B, C, res, flags, stream,
                        TrainOut{x1, alpha, alpha_prev, x_alpha, tar1, tar2});
}

int bndm_get_noise_shard_f32(bndm_L *h, const float *z_global, const float *gamma, float *out, float *out_bn, float *out_wn,
                             int B, int C, int res, unsigned flags, int B_global, int b_offset, void *stream) {
  if (!out) { set_error("bndm_get_noise_shard_f32: null argument"); return BNDM_ERR_ARG; }
  if (B_global < 1) { set_error("bndm_get_noise_shard_f32: bad global batch %d", B_global); return BNDM_ERR_ARG; }
  return get_noise_impl(h, z_global, gamma, out, out_bn, out_wn, B, C, res, flags, stream,
                        TrainOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, B_global, b_offset);
}

int bndm_white128_reinterpret_f32(const float *x, float *out, int B, int C, void *stream) {
  return bndm_white128_reinterpret_shard_f32(x, out, B, C, B, 0, stream);
}

int bndm_white128_reinterpret_shard_f32(const float *x_global, float *out, int B, int C, int B_global, int b_offset, void *stream) {
  if (!x_global || !out || B < 1 || C < 1) { set_error("bndm_white128_reinterpret_f32: bad argument"); return BNDM_ERR_ARG; }
  if (b_offset < 0 || b_offset + B > B_global) { set_error("bndm_white128_reinterpret_shard_f32: shard outside the global batch"); return BNDM_ERR_ARG; }
  if (x_global == out) { set_error("bndm_white128_reinterpret_f32: in-place not supported"); return BNDM_ERR_ARG; }
  CK(launch_white128(x_global, out, B, C, B_global, b_offset, (cudaStream_t)stream));
  return BNDM_OK;
}

static int check_step(const void *x_out, const void *x, const void *d, int B, int C, int HW, int Cd) {
  if (!x_out || !x || !d) { set_error("iadb step: null argument"); return BNDM_ERR_ARG; }
  if (B < 1 || C < 1 || HW < 1) { set_error("iadb step: bad shape"); return BNDM_ERR_ARG; }
  if (Cd != C && Cd != 2 * C) {
    set_error("iadb step: UNet output has %d channels, expected %d or %d", Cd, C, 2 * C);
    return BNDM_ERR_UNSUPPORTED;   // the reference's NotImplementedError (iadb_bn.py:331)
  }
  return BNDM_OK;
}

int bndm_iadb_step_f32(float *x_out, const float *x, const float *d, const float *dalpha, const float *dgamma, int B,
                       int C, int HW, int d_channels, void *stream) {
  int rc = check_step(x_out, x, d, B, C, HW, d_channels);
  if (rc != BNDM_OK) return rc;
  if (!dalpha || (d_channels == 2 * C && !dgamma)) { set_error("iadb step: missing coefficient vector"); return BNDM_ERR_ARG; }
  IadbArgs a{x_out, x, d, dalpha, dgamma, nullptr, nullptr, nullptr, B, C, HW, d_channels, 0};
  CK(launch_iadb_step(a, false, (cudaStream_t)stream));
  return BNDM_OK;
}

static int iadb_sched(float *x_out, const float *x, const float *d, const float *table, int *state, float *t_next_out, int B,
                      int C, int HW, int d_channels, int d_nhwc, void *stream) {
  int rc = check_step(x_out, x, d, B, C, HW, d_channels);
  if (rc != BNDM_OK) return rc;
  if (!table || !state) { set_error("iadb sched step: null table/state"); return BNDM_ERR_ARG; }
  IadbArgs a{x_out, x, d, nullptr, nullptr, table, state, t_next_out, B, C, HW, d_channels, d_nhwc};
  CK(launch_iadb_step(a, true, (cudaStream_t)stream));
  return BNDM_OK;
}

int bndm_iadb_step_sched_f32(float *x_out, const float *x, const float *d, const float *table, int *state,
                             float *t_next_out, int B, int C, int HW, int d_channels, void *stream) {
  return iadb_sched(x_out, x, d, table, state, t_next_out, B, C, HW, d_channels, 0, stream);
}

int bndm_iadb_step_sched_dnhwc_f32(float *x_out, const float *x, const float *d_nhwc, const float *table, int *state,
                                   float *t_next_out, int B, int C, int HW, int d_channels, void *stream) {
  return iadb_sched(x_out, x, d_nhwc, table, state, t_next_out, B, C, HW, d_channels, 1, stream);
}

int bndm_ddim_step_f32(float *x_out, const float *x, const float *eps, const float *noise, const float *coef, int *state,
                       float *t_next_out, int B, int clip, int64_t n, void *stream) {
  if (!x_out || !x || !eps || !coef || n < 1) { set_error("ddim step: bad argument"); return BNDM_ERR_ARG; }
  DdimArgs a{x_out, x, eps, noise, coef, state, t_next_out, B, clip, n};
  CK(launch_ddim_step(a, (cudaStream_t)stream));
  return BNDM_OK;
}

// Host-only self-check of the stream-K schedule (used by the CPU test-suite): every stage is
// covered exactly once, segment slots are unique and below n_slots(), and the slot range the
// combine kernel derives for a row tile is exactly the set of segments the GEMM writes for it.
int bndm_debug_streamk_check(int n_tiles, int dense, int n_colblk, int num_sms) {
  return bndm_debug_streamk_check_sub(n_tiles, dense, n_colblk, num_sms, 1);
}

int bndm_debug_streamk_check_sub(int n_tiles, int dense, int n_colblk, int num_sms, int sub) {
  if (sub != 1 && sub != 2) { set_error("streamk: sub must be 1 or 2"); return BNDM_ERR_ARG; }
  const StreamK k = make_streamk(n_tiles, dense, n_colblk, num_sms, sub);
  if (k.G < 1 || k.G > num_sms || k.W != n_colblk * k.Stot) { set_error("streamk: bad G/W"); return BNDM_ERR_ARG; }
  const int ns = k.n_slots();
  int *owner = new int[ns];
  int *tile_of = new int[ns];
  for (int i = 0; i < ns; ++i) owner[i] = tile_of[i] = -1;
  int rc = BNDM_OK;
  int covered = 0;
  for (int c = 0; c < k.G && rc == BNDM_OK; ++c) {
    const int b = k.cta_begin(c), e = k.cta_begin(c + 1);
    if (e <= b) { set_error("streamk: empty range for cta %d", c); rc = BNDM_ERR_ARG; break; }
    if (c == 0 && b != 0) { set_error("streamk: range does not start at 0"); rc = BNDM_ERR_ARG; break; }
    for (int g = b; g < e;) {
      int cb, tile, s0;
      k.decode(g, cb, tile, s0);
      if (cb < 0 || cb >= n_colblk || tile < 0 || tile >= n_tiles || k.tile_begin(cb, tile) + s0 != g) {
        set_error("streamk: decode(%d) inconsistent", g); rc = BNDM_ERR_ARG; break;
      }
      const int se = e < k.tile_end(cb, tile) ? e : k.tile_end(cb, tile);
      if (k.cta_of(g) != c || k.cta_of(se - 1) != c) { set_error("streamk: cta_of mismatch at %d", g); rc = BNDM_ERR_ARG; break; }
      const int sl = k.slot(c, cb, tile);
      if (sl < 0 || sl >= ns || owner[sl] != -1) { set_error("streamk: slot %d reused/out of range", sl); rc = BNDM_ERR_ARG; break; }
      owner[sl] = c;
      tile_of[sl] = cb * n_tiles + tile;
      covered += se - g;
      g = se;
    }
  }
  if (rc == BNDM_OK && (covered != k.W || k.cta_begin(k.G) != k.W)) { set_error("streamk: covered %d of %d", covered, k.W); rc = BNDM_ERR_ARG; }
  for (int cb = 0; cb < n_colblk && rc == BNDM_OK; ++cb)
    for (int t = 0; t < n_tiles && rc == BNDM_OK; ++t) {
      const int c0 = k.cta_of(k.tile_begin(cb, t)), c1 = k.cta_of(k.tile_end(cb, t) - 1);
      int count = 0;
      for (int sl = 0; sl < ns; ++sl) count += tile_of[sl] == cb * n_tiles + t;
      if (count != c1 - c0 + 1) { set_error("streamk: tile (%d,%d) has %d segments, combine expects %d", cb, t, count, c1 - c0 + 1); rc = BNDM_ERR_ARG; }
      for (int c = c0; c <= c1 && rc == BNDM_OK; ++c)
        if (tile_of[k.slot(c, cb, t)] != cb * n_tiles + t || owner[k.slot(c, cb, t)] != c) {
          set_error("streamk: slot of cta %d for tile (%d,%d) not written by it", c, cb, t); rc = BNDM_ERR_ARG;
        }
    }
  delete[] owner;
  delete[] tile_of;
  return rc;
}

// Host-only check of K1g's row schedule (CPU test-suite): every needed quad is owned by exactly one CTA slot,
// every row group is sorted longest first (the kernel's "active quads are a prefix" rule), and the heaviest CTA's
// load (in pipeline-stage chunks) is reported next to the total so the test can bound the imbalance.
int bndm_debug_gemv_schedule_check(int res32, int dense, int n_ctas, int variant, int *max_load, int *total_load) {
  if (n_ctas < 1 || variant != 0) { set_error("gemv schedule: bad argument"); return BNDM_ERR_ARG; }
  const int kw = gemv_kw(variant);
  int *table = new (std::nothrow) int[(size_t)n_ctas * kGemvTableStride];
  if (!table) { set_error("out of host memory"); return BNDM_ERR_ARG; }
  int rc = BNDM_OK;
  const long blocks = gemv_build_schedule(res32, dense, n_ctas, variant, table);
  if (blocks < 0) { set_error("gemv schedule: quads do not fit %d CTAs", n_ctas); rc = BNDM_ERR_UNSUPPORTED; }
  int seen[kNPix / 4] = {0};
  int worst = 0, total = 0;
  for (int c = 0; c < n_ctas && rc == BNDM_OK; ++c) {
    int load = 0;
    if (table[c * kGemvTableStride + kGemvSlots] != total) { set_error("gemv schedule: stream offset of cta %d", c); rc = BNDM_ERR_ARG; break; }
    for (int sl = 0; sl < kGemvSlots; ++sl) {
      const int q = table[c * kGemvTableStride + sl];
      if (q < 0) {           // empty slots come last (the kernel's "active slots are a prefix" rule)
        for (int r = sl + 1; r < kGemvSlots; ++r)
          if (table[c * kGemvTableStride + r] >= 0) { set_error("gemv schedule: hole in the slots of cta %d", c); rc = BNDM_ERR_ARG; }
        break;
      }
      if (q >= kNPix / 4 || seen[q]++) { set_error("gemv schedule: quad %d out of range or owned twice", q); rc = BNDM_ERR_ARG; break; }
      const int kend = dense ? kNPix : 4 * q + 4;
      load += (kend + kw - 1) / kw;
      if (sl > 0 && !dense && table[c * kGemvTableStride + sl - 1] < q) {
        set_error("gemv schedule: slots of cta %d not sorted longest first", c); rc = BNDM_ERR_ARG; break;
      }
    }
    total += load;
    if (load > worst) worst = load;
  }
  if (rc == BNDM_OK && total != blocks) { set_error("gemv schedule: %d blocks counted, %ld reported", total, blocks); rc = BNDM_ERR_ARG; }
  for (int q = 0; q < kNPix / 4 && rc == BNDM_OK; ++q) {
    const bool needed = !res32 || (q < 512 && (q & 15) < 8);
    if ((seen[q] != 0) != needed) { set_error("gemv schedule: quad %d %s", q, needed ? "missing" : "not needed but scheduled"); rc = BNDM_ERR_ARG; }
  }
  delete[] table;
  if (max_load) *max_load = worst;
  if (total_load) *total_load = total;
  return rc;
}

int bndm_upsample2x_nhwc_f32(const float *x, float *y, int B, int H, int W, int C, void *stream) {
  if (!x || !y || B < 1 || H < 1 || W < 1 || C < 4 || C % 4 != 0) { set_error("upsample2x: bad argument"); return BNDM_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 != 0) { set_error("upsample2x: misaligned"); return BNDM_ERR_ARG; }
  CK(launch_upsample2x_nhwc(x, y, B, H, W, C, (cudaStream_t)stream));
  return BNDM_OK;
}

int bndm_attention_small_f32(const float *qkv, float *out, int B, int T, int C, int head_dim, void *stream) {
  if (!qkv || !out || B < 1 || T < 1 || C < 1) { set_error("attention_small: bad argument"); return BNDM_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) % 16 != 0) { set_error("attention_small: misaligned"); return BNDM_ERR_ARG; }
  cudaError_t e = launch_attention_small(qkv, out, B, T, C, head_dim, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) {
    set_error("attention_small: head_dim must be 8 and T <= 64 (got head_dim=%d, T=%d)", head_dim, T);
    return BNDM_ERR_UNSUPPORTED;
  }
  CK(e);
  return BNDM_OK;
}

int bndm_add_bias_nhwc_f32(const float *a, const float *a2, const float *bias_a, const float *b, const float *bias_b, float *out,
                           int64_t n, int C, void *stream) {
  if (!a || !b || !bias_b || !out || n < 1 || C < 4 || C % 4 != 0 || n % C != 0) { set_error("add_bias: bad argument"); return BNDM_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(bias_b) |
       reinterpret_cast<uintptr_t>(bias_a) | reinterpret_cast<uintptr_t>(a2) | reinterpret_cast<uintptr_t>(out)) % 16 != 0) {
    set_error("add_bias: pointers must be 16-byte aligned");
    return BNDM_ERR_ARG;
  }
  CK(launch_add_bias_nhwc(a, a2, bias_a, b, bias_b, out, (size_t)n, C, (cudaStream_t)stream));
  return BNDM_OK;
}

int bndm_groupnorm_nhwc_f32(const float *x, const float *x2, int C1, const float *res, const float *add_bc, int add_bc_stride,
                            const float *weight, const float *bias, float *sum_out, float *y, int B, int C, int HW, int groups,
                            float eps, int apply_silu, void *stream) {
  if (x2 && (C1 < 4 || C1 >= C || C1 % 4 != 0 || reinterpret_cast<uintptr_t>(x2) % 16 != 0)) {
    set_error("groupnorm: bad second source (C1=%d of C=%d)", C1, C);
    return BNDM_ERR_ARG;
  }
  if (add_bc && ((add_bc_stride != 0 && add_bc_stride < C) || add_bc_stride % 4 != 0)) {     // 0 = one row for all samples
    set_error("groupnorm: bad add_bc stride");
    return BNDM_ERR_ARG;
  }
  if (!x || !weight || !bias || !y || B < 1 || C < 1 || HW < 1 || groups < 1) { set_error("groupnorm: bad argument"); return BNDM_ERR_ARG; }
  if (C % groups != 0 || (C / groups) % 4 != 0) {
    set_error("groupnorm: channels per group must be a multiple of 4 (C=%d, groups=%d)", C, groups);
    return BNDM_ERR_UNSUPPORTED;
  }
  uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(weight) |
                 reinterpret_cast<uintptr_t>(bias);
  if (res) al |= reinterpret_cast<uintptr_t>(res);
  if (add_bc) al |= reinterpret_cast<uintptr_t>(add_bc);
  if (sum_out) al |= reinterpret_cast<uintptr_t>(sum_out);
  if (al % 16 != 0) { set_error("groupnorm: pointers must be 16-byte aligned"); return BNDM_ERR_ARG; }
  cudaError_t e = launch_groupnorm_nhwc(x, x2, C1, res, add_bc, add_bc_stride, weight, bias, sum_out, y, B, C, HW, groups, eps, apply_silu, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) { set_error("groupnorm: unsupported shape C=%d groups=%d", C, groups); return BNDM_ERR_UNSUPPORTED; }
  CK(e);
  return BNDM_OK;
}

int bndm_linear_tc_f32(const float *a, const float *w, const float *bias, float *out, int M, int N, int K, void *stream) {
  if (!a || !w || !out || M < 1 || N < 1 || K < 1) { set_error("linear_tc: bad argument"); return BNDM_ERR_ARG; }
  if (K % 32 != 0 || N % 4 != 0) {
    set_error("linear_tc: K must be a multiple of 32 and N of 4 (M=%d N=%d K=%d)", M, N, K);
    return BNDM_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) % 16 != 0) {
    set_error("linear_tc: pointers must be 16-byte aligned");
    return BNDM_ERR_ARG;
  }
  cudaError_t e = launch_linear_tc(a, w, bias, out, M, N, K, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) { set_error("linear_tc: unsupported shape M=%d N=%d K=%d", M, N, K); return BNDM_ERR_UNSUPPORTED; }
  CK(e);
  return BNDM_OK;
}

int bndm_shortcut_residual_tf32(const float *x, const float *x2, int C1, int C2, const float *w, const float *h2, const float *bias,
                                float *out, int64_t M, int N, void *stream) {
  if (!x || !w || !h2 || !out || M < 1 || N < 1 || C1 < 1 || C2 < 0 || (C2 > 0 && !x2)) { set_error("shortcut_residual: bad argument"); return BNDM_ERR_ARG; }
  if (C1 % 32 != 0 || C2 % 32 != 0 || N % 4 != 0) {
    set_error("shortcut_residual: C1 and C2 must be multiples of 32 and N of 4 (C1=%d C2=%d N=%d)", C1, C2, N);
    return BNDM_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(h2) |
       reinterpret_cast<uintptr_t>(out)) % 16 != 0) {
    set_error("shortcut_residual: pointers must be 16-byte aligned");
    return BNDM_ERR_ARG;
  }
  cudaError_t e = launch_shortcut_tc(x, x2, C1, C2, w, h2, bias, out, (long long)M, N, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) { set_error("shortcut_residual: unsupported shape M=%lld N=%d", (long long)M, N); return BNDM_ERR_UNSUPPORTED; }
  CK(e);
  return BNDM_OK;
}

int bndm_conv_in3x3_nhwc_f32(const float *x, const float *w, float *out, int B, int Cin, int H, int W, int Cout, void *stream) {
  if (!x || !w || !out || B < 1 || Cin < 1 || H < 1 || W < 1 || Cout < 1) { set_error("conv_in3x3: bad argument"); return BNDM_ERR_ARG; }
  if (reinterpret_cast<uintptr_t>(out) % 16 != 0) { set_error("conv_in3x3: out must be 16-byte aligned"); return BNDM_ERR_ARG; }
  cudaError_t e = launch_conv_in3x3(x, w, out, B, Cin, H, W, Cout, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) { set_error("conv_in3x3: unsupported shape Cin=%d Cout=%d", Cin, Cout); return BNDM_ERR_UNSUPPORTED; }
  CK(e);
  return BNDM_OK;
}

int bndm_snapshot_uint8_hwc(const float *x, uint8_t *out, int N, int C, int H, int W, const int *final_flags, int final_all,
                            void *stream) {
  if (!x || !out || N < 1 || C < 1 || H < 1 || W < 1) { set_error("snapshot_uint8: bad argument"); return BNDM_ERR_ARG; }
  if ((int64_t)C * H * W >= ((int64_t)1 << 31)) { set_error("snapshot_uint8: image too large"); return BNDM_ERR_ARG; }
  CK(launch_snapshot_u8(x, out, N, C, H * W, final_flags, final_all, (cudaStream_t)stream));
  return BNDM_OK;
}

int bndm_to_uint8_nhwc(const float *x, uint8_t *out, int B, int C, int H, int W, void *stream) {
  if (!x || !out || B < 1 || C < 1 || H < 1 || W < 1) { set_error("to_uint8: bad argument"); return BNDM_ERR_ARG; }
  CK(launch_to_u8(x, out, B, C, H * W, (cudaStream_t)stream));
  return BNDM_OK;
}

}  // extern "C"
