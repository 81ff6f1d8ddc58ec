// FFMA (CUDA-core) GEMM:  C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)
//
// This is the fp32-parity backend (logits within 1e-4 of the PyTorch-fp32 reference cannot be held by
// single-pass TF32/BF16 tensor-core math through 12+ layers) and the backend of the tiny
// precision-critical head GEMM.  The throughput path is gemm_tc.cuh (tcgen05).
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, register-prefetch double buffering.
// A and W are both K-major ("NT"), i.e. exactly nn.Linear's x[M,K] and weight[N,K].
#pragma once
#include "common.cuh"

namespace bofi {

constexpr int kSBM = 128, kSBN = 128, kSBK = 16, kSPad = 4;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TIn* __restrict__ A, int lda, const TIn* __restrict__ W, int ldw,
                 const float* __restrict__ bias, const float* residual, int ldr,
                 TOut* C, int ldc, int M, int N, int K, int relu, const int* live_rows, const int* rows_dev) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  __shared__ __align__(16) float As[2][kSBK][kSBM + kSPad];
  __shared__ __align__(16) float Bs[2][kSBK][kSBN + kSPad];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
  if (rows_dev && m0 >= *rows_dev) return;            // valid-row count on the device (compacted SAIC step)

  // global->register staging: 128 rows x 16 k = 512 quads per operand, 2 per thread
  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256, row = idx >> 2, kq = (idx & 3) * 4;
      int gm = m0 + row, gn = n0 + row;
      ra[i] = (gm < M) ? load4(A + (size_t)gm * lda + k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = (gn < N) ? load4(W + (size_t)gn * ldw + k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256, row = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
      As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
      Bs[buf][kq + 0][row] = rb[i].x; Bs[buf][kq + 1][row] = rb[i].y;
      Bs[buf][kq + 2][row] = rb[i].z; Bs[buf][kq + 3][row] = rb[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  const int nk = K / kSBK;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) fetch((kb + 1) * kSBK);
#pragma unroll
    for (int k = 0; k < kSBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias (+relu) (+residual) ; rows {ty*4.., 64+ty*4..}, cols {tx*4.., 64+tx*4..}
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int gn = n0 + jh * 64 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = gn + j;
        if (n >= N) continue;
        float v = acc[i][jh * 4 + j] + (bias ? bias[n] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (residual) v += residual[(size_t)gm * ldr + n];
        C[(size_t)gm * ldc + n] = from_float<TOut>(v);
      }
    }
  }
}

template <typename TIn, typename TOut>
inline cudaError_t gemm_simt(cudaStream_t s, const TIn* A, int lda, const TIn* W, int ldw, const float* bias,
                             const float* residual, int ldr, TOut* C, int ldc, int M, int N, int K, int relu,
                             const int* live_rows, const int* rows_dev = nullptr) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K % kSBK != 0 || lda % 4 != 0 || ldw % 4 != 0) return cudaErrorInvalidValue;
  dim3 grid((N + kSBN - 1) / kSBN, (M + kSBM - 1) / kSBM);
  launch_k(gemm_simt_kernel<TIn, TOut>, grid, 256, 0, s, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, relu, live_rows, rows_dev);
  return cudaGetLastError();
}

}  // namespace bofi
