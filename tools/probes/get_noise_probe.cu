// Torch-free driver of the C ABI for quick GPU experiments on get_noise (a gpurun call without `import torch`
// costs ~20 s of box time instead of minutes).  Links libbndm_b200.so; variants are chosen with the library's
// env knobs (BNDM_TC_STAGES, BNDM_TC_FUSED, BNDM_TC_RAWL, BNDM_TC_SUB, BNDM_TC_MAX_NB, BNDM_NO_PDL ...).
//
//   make probes && build/get_noise_probe [B=4] [C=3] [reps=9] [flush=0]   (0: write flush, 1: read flush, 2: none = L2-warm)
//
// Prints: max |error| of out / bn / wn against an fp64 host evaluation of  bn = L z,  out = bn (1-g) + z g
// (64x64 branch, BNDM_SRC_IMAGE), the cold whole-call time (graph of 10 x [L2 flush, call] minus the flushes),
// and the contraction kernel's per-CTA timeline from bndm_debug_set_trace.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../include/bndm_b200.h"

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e__ = (x);                                                                    \
    if (e__ != cudaSuccess) {                                                                 \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__));     \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)
#define CB(x)                                                                                 \
  do {                                                                                        \
    int rc__ = (x);                                                                           \
    if (rc__ != 0) {                                                                          \
      fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #x, rc__, bndm_last_error()); \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)(rng_state >> 33);
}
static inline double uni() { return (rnd() + 0.5) / 2147483648.0; }
static inline float gauss() { return (float)(sqrt(-2.0 * log(uni())) * cos(6.283185307179586 * uni())); }

__device__ float g_sink;
__global__ void flush_kernel(float4 *buf, size_t n, int clean) {      // clean: flush with reads (no dirty lines left)
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (clean) acc += buf[i].x;
    else buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (clean && acc == 12345.678f) g_sink = acc;
}
// one tiny grid between the flush and the call: the contraction is launched with programmatic stream
// serialisation and would otherwise start (and stamp its trace) while the flush is still draining
__global__ void spacer_kernel() {}

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4, C = argc > 2 ? atoi(argv[2]) : 3, reps = argc > 3 ? atoi(argv[3]) : 9;
  const int clean = argc > 4 ? atoi(argv[4]) : 0;
  const int n = 4096, N = B * C;
  const size_t img = (size_t)N * n;
  printf("# get_noise probe: B=%d C=%d (N=%d columns), 64x64, gaussianBN, inplace; sm100=%d\n", B, C, N, bndm_device_is_sm100());

  // lower-triangular L with rows of roughly unit norm (like a Cholesky factor of a unit-diagonal covariance)
  std::vector<float> L((size_t)n * n, 0.f), z(img), gamma(B);
  for (int p = 0; p < n; ++p) {
    const float s = 1.0f / sqrtf((float)(p + 1));
    for (int k = 0; k <= p; ++k) L[(size_t)p * n + k] = gauss() * s;
  }
  for (auto &v : z) v = gauss();
  for (auto &g : gamma) g = (float)uni();

  float *dL, *dz, *dg, *dout, *dbn, *dwn;
  float4 *dflush;
  unsigned long long *dtrace;
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&dL, L.size() * 4));
  CK(cudaMalloc(&dz, img * 4));
  CK(cudaMalloc(&dg, B * 4));
  CK(cudaMalloc(&dout, img * 4));
  CK(cudaMalloc(&dbn, img * 4));
  CK(cudaMalloc(&dwn, img * 4));
  CK(cudaMalloc(&dflush, flush_bytes));
  CK(cudaMalloc(&dtrace, 148 * 24 * 8));
  CK(cudaMemcpy(dL, L.data(), L.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dz, z.data(), img * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, gamma.data(), B * 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  bndm_L *h = nullptr;
  CB(bndm_prepare_L(dL, n, N, st, &h));
  printf("# L lower-triangular: %d, workspace %.1f MB\n", bndm_L_is_lower_triangular(h), bndm_workspace_bytes(h) * 1e-6);

  // ---- parity against fp64 on the host
  CB(bndm_get_noise_f32(h, dz, dg, dout, dbn, dwn, B, C, 64, BNDM_SRC_IMAGE, st));
  CK(cudaStreamSynchronize(st));
  std::vector<float> out(img), bn(img), wn(img);
  CK(cudaMemcpy(out.data(), dout, img * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(bn.data(), dbn, img * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(wn.data(), dwn, img * 4, cudaMemcpyDeviceToHost));
  double e_out = 0, e_bn = 0, e_wn = 0, rms = 0;
  const int check_cols = std::min(N, 24);                       // the first columns and the last one
  for (int jj = 0; jj <= check_cols; ++jj) {
    const int j = jj < check_cols ? jj : N - 1;
    const float g = gamma[j / C];
    for (int p = 0; p < n; ++p) {
      double acc = 0;
      const float *Lr = &L[(size_t)p * n], *zc = &z[(size_t)j * n];
      for (int k = 0; k <= p; ++k) acc += (double)Lr[k] * zc[k];
      const size_t o = (size_t)j * n + p;
      e_bn = std::max(e_bn, fabs(bn[o] - acc));
      e_out = std::max(e_out, fabs(out[o] - (acc * (1.0 - g) + (double)zc[p] * g)));
      e_wn = std::max(e_wn, (double)fabsf(wn[o] - zc[p]));
      rms += acc * acc;
    }
  }
  rms = sqrt(rms / ((double)(check_cols + 1) * n));
  printf("parity vs fp64 host (%d columns): max|bn err| %.3e  max|out err| %.3e  max|wn err| %.1e  (rms of bn %.3f)  %s\n",
         check_cols + 1, e_bn, e_out, e_wn, rms, (e_bn < 2e-5 && e_out < 2e-5 && e_wn == 0) ? "OK" : "FAIL");

  // ---- cold whole-call time: graph of 10 x [flush, call] minus graph of 10 x [flush]
  auto capture = [&](bool with_call) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 10; ++i) {
      if (clean != 2) flush_kernel<<<148 * 4, 512, 0, st>>>(dflush, flush_bytes / 16, clean);
      spacer_kernel<<<1, 32, 0, st>>>();
      if (with_call) CB(bndm_get_noise_f32(h, dz, dg, dout, nullptr, nullptr, B, C, 64, BNDM_SRC_IMAGE, st));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    return ge;
  };
  cudaGraphExec_t g_call = capture(true), g_flush = capture(false);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto time_graph = [&](cudaGraphExec_t ge) {
    std::vector<float> ms;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaEventRecord(e0, st));
      CK(cudaGraphLaunch(ge, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      float t;
      CK(cudaEventElapsedTime(&t, e0, e1));
      if (r >= 2) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    return ms[ms.size() / 2];
  };
  const float t_call = time_graph(g_call), t_flush = time_graph(g_flush);
  const double us = (t_call - t_flush) * 100.0;
  const double bytes = 4.0 * n * (n + 1) / 2 + 4.0 * img * 2;       // L triangle + z read + out written
  printf("whole call (1 output, %s flush): %.2f us  => %.2f TB/s algorithmic (%.1f MB)\n", clean == 2 ? "no (L2-warm)" : clean ? "clean" : "dirty", us, bytes / us * 1e-6, bytes * 1e-6);

  // ---- per-CTA timeline of the contraction kernel
  CB(bndm_debug_set_trace(h, dtrace));
  std::vector<unsigned long long> tr(148 * 24);
  for (int it = 0; it < 3; ++it) {
    if (clean != 2) flush_kernel<<<148 * 4, 512, 0, st>>>(dflush, flush_bytes / 16, clean);
    CK(cudaMemsetAsync(dtrace, 0, 148 * 24 * 8, st));
    spacer_kernel<<<1, 32, 0, st>>>();
    CB(bndm_get_noise_f32(h, dz, dg, dout, nullptr, nullptr, B, C, 64, BNDM_SRC_IMAGE, st));
    CK(cudaStreamSynchronize(st));
  }
  CK(cudaMemcpy(tr.data(), dtrace, tr.size() * 8, cudaMemcpyDeviceToHost));
  int ctas = 0;
  unsigned long long g0 = ~0ull, g1 = 0;
  double ghz = 0;
  for (int c = 0; c < 148; ++c) {
    const unsigned long long *t = &tr[c * 24];
    if (t[0] == 0) continue;
    ++ctas;
    g0 = std::min(g0, t[0]);
    g1 = std::max(g1, t[7]);
    ghz += (double)(t[6] - t[1]) / (double)(t[7] - t[0]);
  }
  if (ctas == 0) { printf("no trace recorded (SIMT path?)\n"); return 0; }
  ghz /= ctas;
  printf("contraction kernel: %d CTAs, span %.2f us, SM clock ~%.2f GHz\n", ctas, (g1 - g0) * 1e-3, ghz);
  auto phase = [&](const char *name, int a, int b, bool only_fused) {
    double sum = 0, mn = 1e30, mx = 0;
    int cnt = 0;
    for (int c = 0; c < 148; ++c) {
      const unsigned long long *t = &tr[c * 24];
      if (t[0] == 0 || (only_fused && t[13] == 0)) continue;
      const double d = (double)(long long)(t[b] - t[a]) / ghz * 1e-3;
      sum += d; mn = std::min(mn, d); mx = std::max(mx, d); ++cnt;
    }
    if (cnt) printf("   %-46s mean %6.2f  min %6.2f  max %6.2f us  (%d CTAs)\n", name, sum / cnt, mn, mx, cnt);
  };
  auto total = [&](const char *name, int k) {
    double sum = 0;
    for (int c = 0; c < 148; ++c) sum += (double)tr[c * 24 + k] / ghz * 1e-3;
    printf("   %-46s mean %6.2f us per CTA\n", name, sum / ctas);
  };
  phase("init (barriers, TMEM alloc, sync)", 1, 2, false);
  phase("init -> first operands landed", 2, 3, false);
  phase("first operands -> last MMA issued", 3, 4, false);
  phase("last MMA issued -> epilogue done", 4, 5, false);
  phase("epilogue done -> exit", 5, 6, false);
  total("producer: waiting for a free stage", 20);
  total("issuer: waiting for operands", 18);
  total("issuer: waiting for a drained TMEM buffer", 19);
  total("issuer: operand wait + MMA issue + commit", 21);
  total("converter: waiting for TMA", 17);
  total("converter: converting (incl. fence + arrive)", 16);
  phase("fused: partial written -> fence done", 8, 9, true);
  phase("fused: fence -> ticket known", 9, 10, true);
  phase("fused: ticket -> acquire fence done", 10, 11, true);
  phase("fused: partial loads + adds (last batch)", 11, 12, true);
  phase("fused: emit (last batch)", 12, 13, true);
  phase("fused: whole combine", 10, 13, true);
  CB(bndm_free_L(h));
  return 0;
}
