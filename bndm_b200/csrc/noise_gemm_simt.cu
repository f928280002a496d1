// K1b (SIMT variant): fp32 FFMA triangular contraction  P[unit][j][r] = sum_k L[128 i + r][k] z[j][k]
// over the unit's k range.  Correct-first kernel that pins every addressing decision and
// stays selectable (BNDM_GEMM_SIMT) as the on-device fp32 reference for the tcgen05 kernel.
// Replaces torch.matmul(cov_mat_L, noise) (get_noise_recent.py:88,113,146).
#include "common.cuh"

namespace bndm {

constexpr int BM = 128, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  int i, kb0, nkb;
  a.sched.decode(blockIdx.x, i, kb0, nkb);
  const int col0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // 8 rows each
  const int ty = tid >> 4;   // 4 cols each

  float acc[4][8];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[n][m] = 0.0f;

  const int k_begin = kb0 * kBlk, k_end = (kb0 + nkb) * kBlk;
  const float *Lrow = a.L + (int64_t)i * kBlk * kNPix;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // A tile: 128 rows x 16 k  (512 float4, 2 per thread), stored k-major
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int f = tid + t * 256;
      const int row = f >> 2, kq = (f & 3) << 2;
      const float4 v = *reinterpret_cast<const float4 *>(Lrow + (int64_t)row * kNPix + k0 + kq);
      As[kq + 0][row] = v.x;
      As[kq + 1][row] = v.y;
      As[kq + 2][row] = v.z;
      As[kq + 3][row] = v.w;
    }
    {  // B tile: 64 columns x 16 k (256 float4)
      const int col = tid >> 2, kq = (tid & 3) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + col < a.n_cols_pad) v = *reinterpret_cast<const float4 *>(a.z + (int64_t)(col0 + col) * kNPix + k0 + kq);
      Bs[kq + 0][col] = v.x;
      Bs[kq + 1][col] = v.y;
      Bs[kq + 2][col] = v.z;
      Bs[kq + 3][col] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][tx * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][tx * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[n][m] = fmaf(av[m], bv[n], acc[n][m]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int col = col0 + ty * 4 + n;
    if (col >= a.n_cols_pad) continue;
    float *P = a.partials + ((int64_t)blockIdx.x * a.n_cols_pad + col) * kBlk + tx * 8;
    *reinterpret_cast<float4 *>(P) = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
    *reinterpret_cast<float4 *>(P + 4) = make_float4(acc[n][4], acc[n][5], acc[n][6], acc[n][7]);
  }
}

cudaError_t launch_gemm_simt(const GemmArgs &a, cudaStream_t s) {
  dim3 grid(a.sched.n_units(), (a.n_cols_pad + BN - 1) / BN);
  gemm_simt_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace bndm
