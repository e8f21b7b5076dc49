// Shared epilogue of the tcgen05 GEMM kernels (gemm_tc.cu: persistent large-M kernel; gemm_skinny.cu: cluster
// split-K kernel for the decode step): bias / exact-erf GELU / tanh(gate) / residual / K/V-cache addressing / store.
#pragma once
#include "common.cuh"

namespace wf {

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct TcEpilogue {
  void* C;
  long long ldc;
  const float* bias;
  const void* residual;
  long long ldr;
  int res_row_mod;
  const float* gate;
  int act;
  int out_f32;
  const int* c_off_ptr;
  long long c_off_mul;
  // head-major output (K/V caches): element (m, n) -> ((m / hm_rpb) * hm_heads + n / 64) * hm_T + m % hm_rpb) * 64 + n % 64
  int hm_heads, hm_T, hm_rpb;
  // split-K: S k-slices per tile; fp32 partials + arrival counters in `ws`, the last-arriving CTA finishes the tile
  int splits;
  float* ws_part;
  int* ws_count;
  // fused LayerNorm of the A operand (gemm_skinny.cu only): the GEMM runs on the RAW rows x with W' = W * diag(gamma);
  // the epilogue applies  y = rstd * (acc - mean * ln_colsum[n]) + bias'[n]  with per-row mean / rstd computed inside
  // the kernel, ln_colsum[n] = sum_k W'[n, k] and bias' = bias + W beta folded into `bias` by the caller.
  const float* ln_colsum;
  float ln_eps;
  // two-destination output (fused q | k,v projection): columns [0, split_n) go row-major to C (ldc, no offset),
  // columns [split_n, N) to C2 with the head-major / offset addressing above (column index rebased to n - split_n)
  int split_n;
  void* C2;
  // LayerNorm statistics across kernels (gemm_tc.cu, M > 128): a GEMM that PRODUCES a residual stream also emits, per
  // output row and per (n-tile, column half), the sum and sum of squares of the values it stored:
  // stat_out[(m * 2 * n_tiles + 2 * n_blk + half) * 2 + {0, 1}]; the GEMM that CONSUMES the stream (ln_colsum set) adds the
  // stat_in_slots partials of its rows up instead of reading the activations a second time.
  float* stat_out;
  const float* stat_in;
  int stat_in_slots;
};

// gemm_skinny.cu
bool skinny_plan(int M, int N, int K, int tile_hint, int* bn_out, int* cs_out, int* bm_out);
int linear_bf16_skinny(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                       const TcEpilogue& ep, int bm, int bn, int cs, cudaStream_t stream);

// gemm_tc2.cu
bool pair_gemm_usable(int M, int N, int K);
int linear_bf16_pair(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                     const TcEpilogue& ep, cudaStream_t stream);

// bias / GELU / gate / residual / convert / store for W (8, 16 or 32) consecutive columns of one output row
template <int W>
__device__ __forceinline__ void finish_chunk(float (&v)[W], const TcEpilogue& ep, int m, long long res_row, int n0,
                                             int N, float gate, long long c_off, float ln_mean = 0.f,
                                             float ln_rstd = 1.f, float2* st = nullptr) {
  const bool full = (n0 + W <= N);
  // bias / colsum are the same for every row: 16-byte loads (a warp-uniform LDG.128 costs one LSU slot for 4 columns;
  // per-element LDG.32 made a plain bias add cost 10 % of the 192000 x 5120 x 1280 GEMM)
  const bool vec = full && ((reinterpret_cast<uintptr_t>(ep.bias) | reinterpret_cast<uintptr_t>(ep.ln_colsum)) & 15) == 0;
  if (ep.ln_colsum) {
    // y = rstd * (acc - mean * colsum[n]) + bias[n]  as two FMAs per element
    const float nm = -ln_rstd * ln_mean;
    if (vec) {
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + n0 + j));
        const float4 b = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j + 0] = fmaf(ln_rstd, v[j + 0], fmaf(nm, cs.x, b.x));
        v[j + 1] = fmaf(ln_rstd, v[j + 1], fmaf(nm, cs.y, b.y));
        v[j + 2] = fmaf(ln_rstd, v[j + 2], fmaf(nm, cs.z, b.z));
        v[j + 3] = fmaf(ln_rstd, v[j + 3], fmaf(nm, cs.w, b.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const bool ok = full || n0 + j < N;
        const float b = (ep.bias && ok) ? __ldg(ep.bias + n0 + j) : 0.f;
        const float cs = ok ? __ldg(ep.ln_colsum + n0 + j) : 0.f;
        v[j] = fmaf(ln_rstd, v[j], fmaf(nm, cs, b));
      }
    }
  } else if (ep.bias) {
    if (vec) {
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
        v[j + 0] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] += (full || n0 + j < N) ? __ldg(ep.bias + n0 + j) : 0.f;
    }
  }
  if (ep.act == 1) {
#pragma unroll
    for (int j = 0; j < W; j += 2) gelu_fast_pair(v[j], v[j + 1]);
  }
  if (ep.gate) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] *= gate;
  }
  long long off;
  void* cbase = ep.C;
  const bool second = ep.split_n > 0 && n0 >= ep.split_n;
  if (ep.hm_heads > 0 && (ep.split_n == 0 || second)) {
    const int nn = n0 - ep.split_n;
    const int b = m / ep.hm_rpb, t = m - b * ep.hm_rpb;
    off = (static_cast<long long>(b * ep.hm_heads + (nn >> 6)) * ep.hm_T + t) * 64 + (nn & 63);
    if (second) cbase = ep.C2;
  } else {
    off = static_cast<long long>(m) * ep.ldc + n0;
    if (ep.split_n > 0) c_off = 0;
  }
  if (ep.out_f32) {
    float* crow = reinterpret_cast<float*>(cbase) + c_off + off;
    if (ep.residual) {
      const float* rrow = reinterpret_cast<const float*>(ep.residual) + res_row * ep.ldr + n0;
#pragma unroll
      for (int j = 0; j < W; ++j) if (full || n0 + j < N) v[j] += rrow[j];
    }
    if (st) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (full || n0 + j < N) { st->x += v[j]; st->y = fmaf(v[j], v[j], st->y); }
    }
    if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < W; j += 4)
        *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (n0 + j < N) crow[j] = v[j];
    }
  } else {
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(cbase) + c_off + off;
    if (ep.residual) {
      const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + res_row * ep.ldr + n0;
      if (full && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < W; j += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(rrow + j);
          v[j + 0] += bf16lo(u.x); v[j + 1] += bf16hi(u.x);
          v[j + 2] += bf16lo(u.y); v[j + 3] += bf16hi(u.y);
          v[j + 4] += bf16lo(u.z); v[j + 5] += bf16hi(u.z);
          v[j + 6] += bf16lo(u.w); v[j + 7] += bf16hi(u.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) if (n0 + j < N) v[j] += __bfloat162float(rrow[j]);
      }
    }
    if (st) {  // statistics of the ROUNDED values: that is what the consumer's LayerNorm will see
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (full || n0 + j < N) {
          const float r = __bfloat162float(__float2bfloat16_rn(v[j]));
          st->x += r;
          st->y = fmaf(r, r, st->y);
        }
    }
    if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < W; j += 8) {
        uint4 u;
        u.x = pack_bf16(v[j + 0], v[j + 1]);
        u.y = pack_bf16(v[j + 2], v[j + 3]);
        u.z = pack_bf16(v[j + 4], v[j + 5]);
        u.w = pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(crow + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) if (n0 + j < N) crow[j] = __float2bfloat16_rn(v[j]);
    }
  }
}


}  // namespace wf
