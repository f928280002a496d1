/* bndm_b200.h -- C ABI of libbndm_b200.so (sm_100a).
 *
 * The reference (xchhuang/bndm) is pure Python: it has no FFI/plugin layer of its own, its
 * boundary for this path is "Python function call, torch tensors in/out" (SURVEY.md 8b).
 * This header is the C-ABI a Python (ctypes) / C++ host binds instead of the torch op
 * sequences cited per entry point.  Plain pointers and sizes only: no torch / C++ types.
 *
 * Conventions
 *   - every pointer marked "dev" is a device pointer on the CURRENT CUDA device, fp32,
 *     contiguous NCHW, 16-byte aligned (torch allocations are 512-byte aligned);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all calls are stream-ordered, never synchronise, never allocate inside a
 *     noise/step call once the handle has enough workspace (=> CUDA-Graph capturable);
 *   - return value: 0 = ok, negative = error (see enum); bndm_last_error() gives the text
 *     (thread-local); nothing throws or exits across the ABI;
 *   - the caller owns every buffer; the library retains no caller pointer past a call,
 *     except the L pointer bound into the opaque bndm_L handle (must outlive it);
 *   - a bndm_L handle owns ONE workspace (gathered columns, operand copies, partial tiles): calls on the same
 *     handle are serialised -- host threads by a mutex inside the handle, streams by an event the library inserts
 *     when a call arrives on another stream than the previous one (a capturing stream relies on the capture
 *     protocol's own ordering; a call on a second stream while the first is still being captured is refused with
 *     BNDM_ERR_WORKSPACE).  Use one handle per stream for concurrent calls; handles are per device.
 */
#ifndef BNDM_B200_H_
#define BNDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNDM_ABI_VERSION 2

enum {
  BNDM_OK = 0,
  BNDM_ERR_ARG = -1,          /* null pointer / bad size / misaligned                          */
  BNDM_ERR_UNSUPPORTED = -2,  /* resolution other than 32/64/128 etc. (the reference's
                                 NotImplementedError sites, get_noise_recent.py:58,166,187)   */
  BNDM_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed                             */
  BNDM_ERR_WORKSPACE = -4,    /* workspace too small while the stream is capturing             */
  BNDM_ERR_ARCH = -5          /* device is not sm_100 (tcgen05 path requested)                 */
};

/* ---- flags for bndm_get_noise_f32 ------------------------------------------------------- */
/* Where the white field comes from (get_noise_recent.py `inplace`):                          */
#define BNDM_SRC_DRAW     0u  /* z is the torch.randn draw the reference would make:
                                 res 64: (B,C,64,64) :108 | res 32: (B,C,64,64) :83 (drawn
                                 AFTER 2x2 tiling) | res 128: (4B,C,64,64) :138                */
#define BNDM_SRC_IMAGE    1u  /* z is the caller's image x (`inplace=True`): (B,C,res,res);
                                 res 32 is tiled 2x2 (:78-79), res 128 is cut in quadrants
                                 concatenated on dim 0 (:131-132)                              */
/* Which contraction kernel:                                                                 */
#define BNDM_GEMM_AUTO    0u  /* default: K1g for <= 16 GEMM columns (HBM-bound GEMV regime),
                                 K1b (tcgen05) above                                          */
#define BNDM_GEMM_SIMT    16u /* fp32 FFMA split-K witness kernel (same results to ~1e-6)     */
#define BNDM_FORCE_DENSE  32u /* ignore the triangular structure of L (testing)               */
#define BNDM_GEMM_GEMV    64u /* force K1g: TMA-streamed fp32 FFMA kernel, complete rows per
                                 CTA, one launch, no split-K (<= 16 columns, else UNSUPPORTED) */
#define BNDM_GEMM_TC      128u /* force K1b: tcgen05 / TMA, error-compensated 3xTF32, stream-K */

typedef struct bndm_L bndm_L; /* opaque: L + its tcgen05 operand copies + workspace           */

int bndm_version(void);
const char *bndm_last_error(void);

/* Device capability probe: 1 if the current device can run the tcgen05 path (cc 10.x).      */
int bndm_device_is_sm100(void);

/* Binds the Cholesky factor `cov_mat_L` (iadb_bn.py:83-86; (n,n) fp32 row-major,
 * L[p_out][p_in], n must be 4096) and prepares what the kernels need: triangularity check
 * (README.md:33 says lower-triangular; verified, dense path otherwise), TF32 hi/lo operand
 * copies for the tcgen05 path, TMA descriptors, workspace for `max_columns` GEMM columns
 * (columns = B*C*(res==128 ? 4 : 1)).  Synchronises `stream` once (init-time call).        */
int bndm_prepare_L(const float *L_dev, int n, int max_columns, void *stream, bndm_L **out);
/* Grow the workspace (not capturable). */
int bndm_reserve_columns(bndm_L *h, int max_columns, void *stream);
int bndm_L_is_lower_triangular(const bndm_L *h);
int64_t bndm_workspace_bytes(const bndm_L *h);
int bndm_free_L(bndm_L *h);

/* Measurement hook (bench.py roofline): when enabled, bndm_get_noise_f32 brackets its three
 * launches (K1a pack, K1b contraction, K1c epilogue) with CUDA events on the launch stream;
 * bndm_profile_last_ms synchronises on the last event and returns the three durations.
 * Not for use under stream capture.                                                        */
int bndm_profile_enable(bndm_L *h, int on);
int bndm_profile_last_ms(bndm_L *h, float *pack_ms, float *gemm_ms, float *epilogue_ms);

/* get_noise_v2, noise_type in {gaussianBN, gaussianRN, GBN}
 * (bluenoise/get_noise_recent.py:73-164; replaces clone + view/permute + torch.matmul
 * (expand+bmm) + permute/contiguous + 4 element-wise kernels + the cat/slice tiling):
 *     bn[b,c,:] = L @ white[b,c,:]   per 64x64 tile
 *     out       = bn*(1-gamma[b]) + wn*gamma[b]      (gamma == NULL  =>  'GBN': out = bn)
 * z      dev  white source, layout per BNDM_SRC_* above
 * gamma  dev  [B] white fraction per sample (the reference's `alpha_t`), or NULL
 * out    dev  (B,C,res,res)            required
 * out_bn dev  (B,C,res,res) or NULL    the reference's `noise_bn`
 * out_wn dev  (B,C,res,res) or NULL    the reference's `noise_wn` (incl. the 128^2
 *                                      pixel/channel re-interpretation, :143-144)
 * out/out_bn/out_wn must not alias z.                                                      */
int bndm_get_noise_f32(bndm_L *h, const float *z, const float *gamma, float *out, float *out_bn,
                       float *out_wn, int B, int C, int res, unsigned flags, void *stream);

/* The same call for ONE SHARD of a batch that is split across GPUs: `z_global` is the white field of
 * the WHOLE batch (B_global samples; layout per BNDM_SRC_* with B = B_global), the outputs and gamma cover samples
 * [b_offset, b_offset + B) only.  Needed because the 128^2 `inplace=True` branch mixes samples across the batch
 * (get_noise_recent.py:131-146: tile n = k*B_global + b is re-read as (b', k') = divmod(n, 4)), so a shard's outputs
 * depend on other shards' white quadrants; for every other branch this equals the unsharded call on the slice.
 * Results are bit-identical to rows [b_offset, b_offset + B) of the unsharded call.                              */
int bndm_get_noise_shard_f32(bndm_L *h, const float *z_global, const float *gamma, float *out, float *out_bn,
                             float *out_wn, int B, int C, int res, unsigned flags, int B_global, int b_offset,
                             void *stream);

/* Training-side fusion (iadb_bn.py:881-954, latent_iadb_bn_diffusers.py:606-633): the
 * same contraction, with the epilogue emitting what the training step builds from get_noise_v2's
 * three results instead of (or next to) them:
 *     x0      = bn*(1-gamma[b]) + wn*gamma[b]             (gamma == NULL: x0 = bn, 'GBN')
 *     x_alpha = alpha[b]*x0 + (1-alpha[b])*x1             iadb_bn.py:915  (x1 = data batch, dev (B,C,res,res))
 *     tar1    = x1 - x0                                   iadb_bn.py:949,976
 *     tar2    = alpha_prev[b]*(bn - wn)                   iadb_bn.py:950  (NULL: not written)
 * alpha, alpha_prev: dev [B].  x0 may be NULL.  z / flags as bndm_get_noise_f32 (the training loop
 * draws: inplace=False => BNDM_SRC_DRAW).  fp32 with the reference's association => bit-identical
 * to the torch expressions applied to this library's (x0, bn, wn).                            */
int bndm_get_noise_train_f32(bndm_L *h, const float *z, const float *gamma, const float *x1, const float *alpha,
                             const float *alpha_prev, float *x_alpha, float *tar1, float *tar2, float *x0,
                             int B, int C, int res, unsigned flags, void *stream);

/* 'gaussian' pass-through at 128^2 with train_or_test=='test' (get_noise_recent.py:50-56):
 * out = noise_padding(reinterpret(quadrants(x))).  x, out: (B,C,128,128), no aliasing.     */
int bndm_white128_reinterpret_f32(const float *x, float *out, int B, int C, void *stream);
/* Same for one shard: x_global is (B_global,C,128,128), out covers samples [b_offset, b_offset + B).           */
int bndm_white128_reinterpret_shard_f32(const float *x_global, float *out, int B, int C, int B_global, int b_offset,
                                        void *stream);

/* IADB update (iadb_bn.py:326,329,344; utils.py:218,221,226; latent...:110,113,117):
 *     x_out = (x + dalpha[b]*d[:, :C]) + dgamma[b]*d[:, C:2C]      (d_channels == 2C)
 *     x_out =  x + dalpha[b]*d                                      (d_channels == C)
 * fp32, no FMA contraction, reference association => bit-exact vs the torch expression.
 * x_out may alias x.  dalpha/dgamma: dev [B]; dgamma may be NULL when d_channels == C.     */
int bndm_iadb_step_f32(float *x_out, const float *x, const float *d, const float *dalpha,
                       const float *dgamma, int B, int C, int HW, int d_channels, void *stream);

/* Same update driven by a device-resident schedule so that ONE captured CUDA graph of
 * [UNet, step] replays for every t.  `table` is [T][B] rows of 4 floats
 * {dalpha, dgamma, t_next, 0}: per step AND per sample, like the (B,) coefficient tensors the
 * reference forms each step (iadb_bn.py:306-316).  Row used for sample b: table[step][b].
 * The kernel also fills t_next_out[b] with its row's t_next (the next UNet "timestep" =
 * alpha_start of the following step, iadb_bn.py:311,319).
 * `state` is 2 ints on the device: state[0] counts block tickets over the run (a launch of G
 * blocks is step ticket / G); state[1] = T, the number of rows of `table` (0 = unchecked): a launch
 * past the table re-uses the last row instead of reading beyond it and sets bit 30 of state[1]
 * (overrun flag).  Before the first step of a run set state = {0, T}; keep B, C, HW, the buffers'
 * alignment and the d layout fixed within a run (they determine G).                          */
int bndm_iadb_step_sched_f32(float *x_out, const float *x, const float *d, const float *table,
                             int *state, float *t_next_out, int B, int C, int HW, int d_channels,
                             void *stream);
/* Same, with the UNet output in channels-last memory (d_nhwc: [B][HW][d_channels]) -- what the
 * channels-last UNet evaluation produces natively, consumed in place (no NCHW copy); x stays NCHW. */
int bndm_iadb_step_sched_dnhwc_f32(float *x_out, const float *x, const float *d_nhwc, const float *table,
                                   int *state, float *t_next_out, int B, int C, int HW, int d_channels,
                                   void *stream);

/* DDIM update, epsilon prediction (diffusers DDIMScheduler.step as called at
 * ddim_diffusers.py:680; parity unpinned, see oracle/sampler.py):
 *     x0  = clamp((x - c[1]*eps) / c[0], -1, 1)          (clamp iff clip != 0)
 *     out = (c[2]*x0 + c[3]*eps) [+ c[4]*noise]          (noise may be NULL => eta = 0)
 * coef: dev, rows of 8 floats {sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev),
 * sqrt(1-abar_prev-sigma^2), sigma, t_next, 0, 0}; the row used is ticket / gridDim like the
 * scheduled IADB step (state = {run-long block-ticket counter, number of rows (0 = unchecked)} as there;
 * state == NULL => row 0).  t_next_out: dev [B] float or NULL.  n = B*C*H*W.  x_out may alias x. */
int bndm_ddim_step_f32(float *x_out, const float *x, const float *eps, const float *noise,
                       const float *coef, int *state, float *t_next_out, int B, int clip, int64_t n,
                       void *stream);

/* Debug / test hook: force the contraction's variants (-1 = default rule, 0 = off, 1 = on):
 * fused_combine = the last CTA of a row tile sums its partial tiles inside the contraction
 * (default: off -- slower than the separate combine launch); raw_L = raw fp32 L blocks + in-kernel lo conversion (default:
 * one column block of <= 64 columns).  Process-wide.                                         */
int bndm_debug_set_policy(int fused_combine, int raw_L);

/* Debug: when `trace_dev` (device, 24 x u64 per CTA, >= 148 CTAs) is non-NULL the tcgen05
 * contraction kernel records per-CTA time stamps {globaltimer in, clock in, after init, first
 * operands landed, last MMA issued, epilogue done, clock out, globaltimer out}.  K1g uses 128 x u64 per CTA:
 * [0] start, [1] first stage landed, [2] stream consumed, [3] outputs stored, [8+c] stage c requested, [48+c] stage c
 * landed (seen by the warp of the longest quad), [88+c] released by it (%globaltimer, c < 40).                    */
int bndm_debug_set_trace(bndm_L *h, unsigned long long *trace_dev);

/* Host-only consistency check of the contraction kernel's stream-K work split (no device
 * work; used by the CPU test-suite).  0 if every pipeline stage is covered exactly once and
 * the combine kernel's view of the partial tiles matches what the GEMM kernel writes.       */
int bndm_debug_streamk_check(int n_tiles, int dense, int n_colblk, int num_sms);
/* Same with `sub` k-stages per schedule unit (1, or 2 = the 64-k pipeline stages of the raw-operand
 * variant).                                                                                    */
int bndm_debug_streamk_check_sub(int n_tiles, int dense, int n_colblk, int num_sms, int sub);

/* Host-only check of K1g's row schedule (no device work; CPU test-suite): 0 if every quad of 4 rows the branch needs
 * (all 1024, or the 256 with h,w < 32 at 32^2) is owned by exactly one CTA and every row group is sorted longest
 * first; *max_load / *total_load = heaviest CTA's / all CTAs' pipeline-stage chunks (load balance).  variant: 0.      */
int bndm_debug_gemv_schedule_check(int res32, int dense, int n_ctas, int variant, int *max_load, int *total_load);

/* K5 -- the UNet's normalisation glue on channels-last activations (diffusers ResnetBlock2D
 * norm/act sequence of the model built at iadb_bn.py:205-282 and called at :319), one kernel:
 *     s = x (+ res) (+ add_bc[b][c]);  sum_out = s (if non-NULL)
 *     (x2 != NULL: the input is the channel concatenation [x | x2] -- x holds channels [0, C1) as
 *      [B][HW][C1], x2 the rest as [B][HW][C - C1] -- i.e. torch.cat((x, x2), 1) without the copy;
 *      res and sum_out must then be NULL)
 *     y = act((s - mean_group) * rstd_group * weight[c] + bias[c]),  act = SiLU iff apply_silu
 * x, res, sum_out, y: dev, NHWC [B][HW][C] fp32; add_bc: dev, B rows of C floats `add_bc_stride`
 * floats apart (a column slice of a wider matrix; 0 = the same row for every sample, i.e. a
 * per-channel bias), or NULL; weight, bias: dev [C].
 * groups as torch.nn.GroupNorm (biased variance, eps inside the sqrt); C/groups % 4 == 0.
 * Replaces RowwiseMoments + affine + SiLU (+ broadcast / residual add) kernels of PyTorch.     */
int bndm_groupnorm_nhwc_f32(const float *x, const float *x2, int C1, const float *res, const float *add_bc,
                            int add_bc_stride, const float *weight, const float *bias, float *sum_out, float *y, int B,
                            int C, int HW, int groups, float eps, int apply_silu, void *stream);

/* K8 -- nearest-neighbour 2x upsampling of an NHWC fp32 activation: x dev [B][H][W][C] -> y dev
 * [B][2H][2W][C] (diffusers Upsample2D's F.interpolate(scale_factor=2, mode="nearest")).  C % 4 == 0. */
int bndm_upsample2x_nhwc_f32(const float *x, float *y, int B, int H, int W, int C, void *stream);

/* K7 -- softmax(q k^T / sqrt(head_dim)) v for the UNet's attention blocks at their tiny sizes (4x4 / 2x2
 * resolution, head_dim 8): qkv dev [B][T][3C] (q | k | v on the last axis, head h = channels 8h..8h+7),
 * out dev [B][T][C].  head_dim must be 8, T <= 64.  Replaces F.scaled_dot_product_attention there. */
int bndm_attention_small_f32(const float *qkv, float *out, int B, int T, int C, int head_dim, void *stream);

/* K9 -- the fp32 linears of the UNet's attention blocks (diffusers Attention to_q/to_k/to_v/to_out; torch.nn.Linear
 * semantics, which the reference runs as fp32 SIMT GEMMs: torch.backends.cuda.matmul.allow_tf32 is False) as a 3xTF32
 * tcgen05 GEMM with fp32-grade results:   out[m][n] = sum_k a[m][k] * w[n][k] (+ bias[n]).
 * a dev [M][K], w dev [N][K] (the Linear's weight), bias dev [N] or NULL, out dev [M][N]; K % 32 == 0, N % 4 == 0,
 * a / w / out 16-byte aligned.                                                                                      */
int bndm_linear_tc_f32(const float *a, const float *w, const float *bias, float *out, int M, int N, int K, void *stream);

/* K10 -- the tail of a ResnetBlock2D with a 1x1 conv_shortcut (diffusers; every resnet of the up blocks, whose input is
 * cat(h, skip)):   out[m][n] = sum_k w[n][k] * cat(x, x2)[m][k]  +  h2[m][n]  +  bias[n]
 * i.e. the shortcut convolution (TF32 inputs, fp32 accumulation, as cuDNN runs it when torch.backends.cudnn.allow_tf32 is
 * set -- the reference's configuration), the residual add and both biases in one tcgen05 kernel.
 * x dev [M][C1], x2 dev [M][C2] or NULL with C2 = 0 (channels-last activations, M = B*H*W rows), w dev [N][C1 + C2] (the
 * convolution's weight), h2 dev [M][N] (conv2's output without its bias), bias dev [N] or NULL (conv_shortcut.bias +
 * conv2.bias), out dev [M][N] (may be h2).  C1 % 32 == 0, C2 % 32 == 0, N % 4 == 0.                                      */
int bndm_shortcut_residual_tf32(const float *x, const float *x2, int C1, int C2, const float *w, const float *h2,
                                const float *bias, float *out, int64_t M, int N, void *stream);

/* K11 -- the UNet's first convolution (diffusers UNet2DModel.conv_in: 3x3, padding 1) from the sampler's NCHW state to the
 * channels-last activation: out[b][h][w][co] = sum w[co][ci][r][s] * x[b][ci][h+r-1][w+s-1], fp32 FMA, no bias.
 * x dev [B][Cin][H][W] (Cin <= 4), w dev [Cout][Cin][3][3] (contiguous), out dev [B][H][W][Cout]; Cout % 4 == 0 and
 * 256 % (Cout / 4) == 0 (128 for every model of the reference), otherwise BNDM_ERR_UNSUPPORTED.                      */
int bndm_conv_in3x3_nhwc_f32(const float *x, const float *w, float *out, int B, int Cin, int H, int W, int Cout, void *stream);

/* K6 -- out = ((a [+ a2]) [+ bias_a[c]]) + (b + bias_b[c]) on NHWC fp32 activations (n elements, C
 * channels innermost; a2 and bias_a may be NULL): the biases of conv_shortcut / conv2 (or an attention block's
 * to_out) and the residual add in one pass, same association as PyTorch's conv-bias then add.
 * out may alias a or b.                                                                        */
int bndm_add_bias_nhwc_f32(const float *a, const float *a2, const float *bias_a, const float *b, const float *bias_b,
                           float *out, int64_t n, int C, void *stream);

/* Image post-processing of the test drivers (iadb_bn.py:796-816, ddim_diffusers.py:687-688):
 * out_u8[b,h,w,c] = round(clamp(x[b,c,h,w]/2 + 0.5, 0, 1) * 255), NCHW fp32 -> NHWC uint8. */
int bndm_to_uint8_nhwc(const float *x, uint8_t *out, int B, int C, int H, int W, void *stream);

/* The IADB test driver's PNG conversion (iadb_bn.py:796-816) for N images x dev (N,C,H,W) -> out dev (N,H,W,C) uint8:
 *   final image (final_flags[n] != 0, or final_all when final_flags == NULL):  v = clamp((x + 1) / 2, 0, 1)
 *   intermediate snapshot:  v = (x - min) / (max - min), min / max over the whole image (:802)
 *   out = (uint8)(v * 255)  -- numpy's astype(uint8), i.e. TRUNCATION (bndm_to_uint8_nhwc rounds, as ddim_diffusers.py does).
 * Replaces 4-6 torch kernels + a D2H copy of the fp32 image per snapshot.  final_flags: dev int[N] or NULL.          */
int bndm_snapshot_uint8_hwc(const float *x, uint8_t *out, int N, int C, int H, int W, const int *final_flags,
                            int final_all, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BNDM_B200_H_ */
