// fp32 linear layer on the CUDA cores:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// The fp32 engine exists for the token-exact parity path (BASELINE config 1: "greedy-decoded token
// IDs identical on the fp32 path"): plain fp32 FMA accumulation, master weights read in place
// (reference whisper/model.py:35-41 with x.dtype == float32, where the per-call cast is a no-op).
// Same epilogue contract as the tcgen05 bf16 kernel in gemm_tc.cu.
#include "common.cuh"
#include "kernels.h"

namespace wf {

static constexpr int FT = 64;   // C tile 64 x 64
static constexpr int FK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw, int M, int N,
                int K, LinearEpilogue ep) {
  __shared__ float As[FK][FT + 4];
  __shared__ float Ws[FK][FT + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FT, n0 = blockIdx.y * FT;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int tx = tid & 15, ty = tid >> 4;  // thread computes rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool w_vec = (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);

  for (int k0 = 0; k0 < K; k0 += FK) {
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, w4[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const int m = m0 + lrow, k = k0 + lk;
      if (m < M) {
        const float* p = A + static_cast<long long>(m) * lda + k;
        if (a_vec && k + 3 < K) {
          const float4 t = *reinterpret_cast<const float4*>(p);
          a4[0] = t.x; a4[1] = t.y; a4[2] = t.z; a4[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < K) a4[i] = p[i];
        }
      }
      const int n = n0 + lrow;
      if (n < N) {
        const float* p = W + static_cast<long long>(n) * ldw + k;
        if (w_vec && k + 3 < K) {
          const float4 t = *reinterpret_cast<const float4*>(p);
          w4[0] = t.x; w4[1] = t.y; w4[2] = t.z; w4[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < K) w4[i] = p[i];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lrow] = a4[i];
      Ws[lk + i][lrow] = w4[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }

  long long c_off = 0;
  if (ep.c_off_ptr) c_off = static_cast<long long>(*ep.c_off_ptr) * ep.c_off_mul;
  const float gate = ep.gate ? tanhf(*ep.gate) : 1.0f;
  float* C = reinterpret_cast<float*>(ep.C) + c_off;
  const float* R = reinterpret_cast<const float*>(ep.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const long long rr = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[n];
      if (ep.act == 1) v = gelu_erf(v);
      if (ep.gate) v *= gate;
      if (R) v += R[rr * ep.ldr + n];
      long long off;
      if (ep.hm_heads > 0) {
        const int b = m / ep.hm_rpb, t = m - b * ep.hm_rpb;
        off = (static_cast<long long>(b * ep.hm_heads + (n >> 6)) * ep.hm_T + t) * 64 + (n & 63);
      } else {
        off = static_cast<long long>(m) * ep.ldc + n;
      }
      C[off] = v;
    }
  }
}

int linear_f32(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K,
               const LinearEpilogue& e, cudaStream_t stream) {
  WF_REQUIRE(M > 0 && N > 0 && K > 0, "linear_f32: empty problem M=%d N=%d K=%d", M, N, K);
  if (e.hm_heads > 0)
    WF_REQUIRE(N % 64 == 0 && e.hm_heads * 64 == N && e.hm_T > 0 && e.hm_rpb > 0 && !e.residual,
               "linear_f32: bad head-major output spec");
  dim3 grid((M + FT - 1) / FT, (N + FT - 1) / FT);
  WF_REQUIRE(grid.y <= 65535, "linear_f32: N=%d too large for this kernel", N);
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(A, lda, W, ldw, M, N, K, e);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

}  // namespace wf
