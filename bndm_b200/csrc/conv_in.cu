// K11: the UNet's first convolution (diffusers UNet2DModel.conv_in: 3x3, padding 1, 3 or 4 input channels -> 128; the
// model the reference samples with, iadb_bn.py:205-282 / :319) from the sampler's NCHW state straight to the channels-last
// activation the fused UNet works on.
//
//     out[b][h][w][co] = sum_{ci, r, s} w[co][ci][r][s] * x[b][ci][h + r - 1][w + s - 1]            (no bias: it is owed, see fused_unet.py)
//
// With 3 input channels cuDNN has no tensor-core kernel for it: it pads the input to 4 channels, converts layouts and runs a
// legacy direct kernel -- 58 + 19 us in five launches per forward at batch 64, plus torch's NCHW -> NHWC copy -- for 1.8 GFLOP
// and one 134 MB write.  Here: fp32 FFMA (27-36 MACs per output, exact fp32 like the rest of the glue), weights staged once
// per persistent CTA in shared memory as [tap][co], a thread owns 4 output channels x 4 neighbouring pixels (16 accumulators:
// 48 FFMA per 9 shared-memory loads), a warp stores 512 contiguous bytes per pixel.  Measured at batch 64, 64^2: 61.7 us in one
// launch against ~81 us in five (cuDNN's kernel + padding / layout helpers + torch's copy); the FFMA floor is 26 us.
#include "common.cuh"

namespace bndm {

constexpr int kCiThreads = 256;
constexpr int kCiMaxCin = 4;
constexpr int kCiTileW = 32;      // pixels of one image row per tile

// kCin / kCout as template constants (0 = take the run-time values): with constant strides the ~190 integer instructions of
// address arithmetic per input channel fold into immediates and the kernel is bound by its 144 FFMA per channel, not by issue
// slots (ncu on the generic instance: IPC 2.75 with half of the instructions integer: 78 us; specialised: 61.7 us)
template <int kCin, int kCout>
__global__ void __launch_bounds__(kCiThreads) conv_in3x3_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                float *__restrict__ out, int B, int Cin_rt, int H, int W, int Cout_rt) {
  extern __shared__ __align__(16) float ci_smem[];
  const int Cin = kCin ? kCin : Cin_rt, Cout = kCout ? kCout : Cout_rt;
  const int cq_n = Cout >> 2;                      // channel quads; kCiThreads % cq_n == 0
  const int pg_n = kCiThreads / cq_n;              // pixel groups of 4 per pass
  const int taps = Cin * 9;
  float4 *sW = reinterpret_cast<float4 *>(ci_smem);                       // [taps][cq_n]
  float *sX = ci_smem + (size_t)taps * Cout;                              // [Cin][3][kCiTileW + 2]
  const int tid = threadIdx.x;
  // weights: w[co][ci][r][s] -> sW[(ci * 9 + r * 3 + s)][co]
  for (int i = tid; i < taps * Cout; i += kCiThreads) {
    const int co = i / taps, t = i - co * taps;
    ci_smem[(size_t)t * Cout + co] = __ldg(w + i);
  }
  const int cq = tid % cq_n, pg = tid / cq_n;
  const int tiles_w = (W + kCiTileW - 1) / kCiTileW;
  const long long n_tiles = (long long)B * H * tiles_w;
  const int sxw = kCiTileW + 2;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const long long bh = tile / tiles_w;
    const int h = (int)(bh % H), b = (int)(bh / H);
    const int w0 = tw * kCiTileW;
    __syncthreads();                               // previous tile's readers are done (and the weights are staged)
    for (int i = tid; i < Cin * 3 * sxw; i += kCiThreads) {
      const int ci = i / (3 * sxw), rem = i - ci * 3 * sxw;
      const int r = rem / sxw, j = rem - r * sxw;
      const int hh = h + r - 1, ww = w0 + j - 1;
      sX[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (((size_t)b * Cin + ci) * H + hh) * W + ww) : 0.0f;
    }
    __syncthreads();
    for (int p0 = pg * 4; p0 < kCiTileW; p0 += pg_n * 4) {
      if (w0 + p0 >= W) break;
      float4 acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ci = 0; ci < (kCin ? kCin : kCiMaxCin); ++ci) {
        if (!kCin && ci >= Cin) break;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float *row = sX + (ci * 3 + r) * sxw + p0;
          float in[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) in[j] = row[j];
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const float4 wv = sW[(size_t)(ci * 9 + r * 3 + s) * cq_n + cq];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j].x = fmaf(wv.x, in[j + s], acc[j].x);
              acc[j].y = fmaf(wv.y, in[j + s], acc[j].y);
              acc[j].z = fmaf(wv.z, in[j + s], acc[j].z);
              acc[j].w = fmaf(wv.w, in[j + s], acc[j].w);
            }
          }
        }
      }
      float4 *dst = reinterpret_cast<float4 *>(out + (((size_t)b * H + h) * W + w0 + p0) * Cout) + cq;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (w0 + p0 + j < W) dst[(size_t)j * cq_n] = acc[j];
    }
  }
}

cudaError_t launch_conv_in3x3(const float *x, const float *w, float *out, int B, int Cin, int H, int W, int Cout, cudaStream_t s) {
  if (B < 1 || Cin < 1 || Cin > kCiMaxCin || H < 1 || W < 1 || Cout < 4 || Cout % 4 != 0 || kCiThreads % (Cout / 4) != 0 || Cout > 4 * kCiThreads)
    return cudaErrorNotSupported;
  const size_t smem = ((size_t)Cin * 9 * Cout + (size_t)Cin * 3 * (kCiTileW + 2)) * sizeof(float);
  if (smem > 96 * 1024) return cudaErrorNotSupported;
  const long long n_tiles = (long long)B * H * ((W + kCiTileW - 1) / kCiTileW);
  long long blocks = 148 * 4;                      // persistent: the weights are staged once per CTA (four CTAs fit on an SM)
  if (blocks > n_tiles) blocks = n_tiles;
  if (Cout == 128 && Cin == 3) conv_in3x3_kernel<3, 128><<<(unsigned)blocks, kCiThreads, smem, s>>>(x, w, out, B, Cin, H, W, Cout);
  else if (Cout == 128 && Cin == 4) conv_in3x3_kernel<4, 128><<<(unsigned)blocks, kCiThreads, smem, s>>>(x, w, out, B, Cin, H, W, Cout);
  else conv_in3x3_kernel<0, 0><<<(unsigned)blocks, kCiThreads, smem, s>>>(x, w, out, B, Cin, H, W, Cout);
  return cudaGetLastError();
}

}  // namespace bndm
