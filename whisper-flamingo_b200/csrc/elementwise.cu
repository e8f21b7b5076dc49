// Row-wise / elementwise kernels around the GEMMs: LayerNorm, conv-stem im2col, token embedding, casts.
// All are HBM-streaming: 16-byte vector loads, one warp per row for LayerNorm (two-pass in registers).
#include "common.cuh"
#include "kernels.h"

namespace wf {

// ------------------------------------------------------------------ LayerNorm (model.py:30-32: fp32 math, eps 1e-5)
// One warp per row; the row is held in registers (d <= 32 * 8 * LN_MAX_VEC) so x is read exactly once.
static constexpr int LN_MAX_VEC = 5;  // 5 * 256 = 1280 elements max per row for the register path

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
  v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ w, const float* __restrict__ b,
                 T* __restrict__ y, long long ldy, int rows, int d, float eps, int vec_ok) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * ldx;
  T* yr = y + row * ldy;
  float v[LN_MAX_VEC][8];
  const int nvec = (d + 255) / 256;  // chunks of 256 elements (8 consecutive per lane)
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAX_VEC; ++c) {
    if (c < nvec) {
      const int i0 = c * 256 + lane * 8;
      if (vec_ok && i0 + 8 <= d) {
        load8(xr + i0, v[c]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[c][j] = (i0 + j < d) ? to_f32(xr[i0 + j]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[c][j];
    }
  }
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAX_VEC; ++c) {
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = c * 256 + lane * 8 + j;
        const float dlt = i < d ? v[c][j] - mean : 0.f;
        sq += dlt * dlt;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / d + eps);
#pragma unroll
  for (int c = 0; c < LN_MAX_VEC; ++c) {
    if (c < nvec) {
      const int i0 = c * 256 + lane * 8;
      if (vec_ok && i0 + 8 <= d) {
        float wv[8], bv[8], o[8];
        load8(w + i0, wv);
        load8(b + i0, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * wv[j] + bv[j];
        store8(yr + i0, o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (i0 + j < d) yr[i0 + j] = from_f32<T>((v[c][j] - mean) * rstd * __ldg(w + i0 + j) + __ldg(b + i0 + j));
      }
    }
  }
}

int layernorm(int dtype, const void* x, long long ldx, const float* w, const float* b, void* y, long long ldy,
              int rows, int d, float eps, cudaStream_t stream) {
  WF_REQUIRE(rows > 0 && d > 0, "layernorm: empty input");
  WF_REQUIRE(d <= LN_MAX_VEC * 256, "layernorm: d=%d exceeds %d", d, LN_MAX_VEC * 256);
  const int warps = 8;
  const int grid = (rows + warps - 1) / warps;
  const auto al = [](const void* p, unsigned a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const int vec_ok = (d % 8 == 0) && (ldx % 8 == 0) && (ldy % 8 == 0) && al(x, 32) && al(y, 32) && al(w, 16) && al(b, 16);
  if (dtype == WF_F32)
    WF_CHECK_CUDA(launch_pdl(1, layernorm_kernel<float>, dim3(grid), dim3(warps * 32), 0, stream, (const float*)x, ldx, w,
                             b, (float*)y, ldy, rows, d, eps, vec_ok));
  else if (dtype == WF_BF16)
    WF_CHECK_CUDA(launch_pdl(1, layernorm_kernel<__nv_bfloat16>, dim3(grid), dim3(warps * 32), 0, stream,
                             (const __nv_bfloat16*)x, ldx, w, b, (__nv_bfloat16*)y, ldy, rows, d, eps, vec_ok));
  else
    WF_REQUIRE(false, "layernorm: bad dtype %d", dtype);
  count_launch();
  return WF_OK;
}

// ------------------------------------------------------------------ im2col for the kernel-3 conv stem
// out[(b*T_out + to) * 3C + c*3 + tap] = in(b, c, to*stride + tap - 1)   (zero outside [0, T_in))
// matches the natural [C_out, C_in, 3] -> [C_out, 3 C_in] flattening of the conv weights (model.py:222-223).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
im2col_k3_kernel(const TI* __restrict__ in, long long sb, long long sc, long long st, int C, int T_in, int T_out,
                 int stride, TO* __restrict__ out, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx enumerates (b, to, c); consecutive threads walk the contiguous input axis for coalescing.
  int c, to;
  long long b;
  if (st == 1) {  // NCW input (mel): t fastest
    to = static_cast<int>(idx % T_out);
    c = static_cast<int>((idx / T_out) % C);
    b = idx / (static_cast<long long>(T_out) * C);
  } else {  // NWC input: c fastest
    c = static_cast<int>(idx % C);
    to = static_cast<int>((idx / C) % T_out);
    b = idx / (static_cast<long long>(T_out) * C);
  }
  const TI* src = in + b * sb + c * sc;
  TO* dst = out + (b * T_out + to) * (3LL * C) + c * 3;
#pragma unroll
  for (int tap = 0; tap < 3; ++tap) {
    const int ti = to * stride + tap - 1;
    const float v = (ti >= 0 && ti < T_in) ? to_f32(src[ti * st]) : 0.f;
    dst[tap] = from_f32<TO>(v);
  }
}

int im2col_k3(int in_dtype, int out_dtype, const void* in, long long sb, long long sc, long long st, int B, int C,
              int T_in, int stride, void* out, cudaStream_t stream) {
  WF_REQUIRE(B > 0 && C > 0 && T_in > 0 && (stride == 1 || stride == 2), "im2col: bad shape");
  const int T_out = (T_in + 2 - 3) / stride + 1;
  const long long total = static_cast<long long>(B) * C * T_out;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (in_dtype == WF_F32 && out_dtype == WF_F32)
    im2col_k3_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)in, sb, sc, st, C, T_in, T_out, stride, (float*)out, total);
  else if (in_dtype == WF_F32 && out_dtype == WF_BF16)
    im2col_k3_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)in, sb, sc, st, C, T_in, T_out, stride, (__nv_bfloat16*)out, total);
  else if (in_dtype == WF_BF16 && out_dtype == WF_BF16)
    im2col_k3_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, sb, sc, st, C, T_in, T_out, stride, (__nv_bfloat16*)out, total);
  else
    WF_REQUIRE(false, "im2col: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

// ------------------------------------------------------------------ token + positional embedding (model.py:306-310)
// out[r * n_pos + j] = T(tok_emb[tokens[r * tok_stride + pos + j]] + pos_emb[pos + j]),  j in [0, n_pos)
// The tables are the fp32 master parameters (read in place); the sum is rounded once to the activation
// dtype like the reference's `.to(xa.dtype)`.  pos comes from device memory when pos_ptr is given, so the
// launch can sit inside a replayed CUDA graph.
template <typename T>
__global__ void __launch_bounds__(128)
embed_kernel(const int* __restrict__ tokens, long long tok_stride, const int* __restrict__ pos_ptr, int pos_const,
             const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, T* __restrict__ out, long long ldo,
             int d, int n_pos) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, j = blockIdx.y;
  const int pos = (pos_ptr ? *pos_ptr : pos_const) + j;
  const int tok = tokens[r * tok_stride + pos];
  const float* te = tok_emb + static_cast<long long>(tok) * d;
  const float* pe = pos_emb + static_cast<long long>(pos) * d;
  T* o = out + (static_cast<long long>(r) * n_pos + j) * ldo;
  for (int i = threadIdx.x; i < d; i += blockDim.x) o[i] = from_f32<T>(te[i] + pe[i]);
}

int embed_tokens(int dtype, const int* tokens, long long tok_stride, const int* pos_ptr, int pos_const, int n_pos,
                 const float* tok_emb, const float* pos_emb, void* out, long long ldo, int R, int d,
                 cudaStream_t stream) {
  WF_REQUIRE(R > 0 && d > 0 && n_pos > 0 && n_pos <= 65535, "embed: bad shape R=%d d=%d n_pos=%d", R, d, n_pos);
  dim3 grid(R, n_pos);
  if (dtype == WF_F32)
    WF_CHECK_CUDA(launch_pdl(3, embed_kernel<float>, grid, dim3(128), 0, stream, tokens, tok_stride, pos_ptr, pos_const,
                             tok_emb, pos_emb, (float*)out, ldo, d, n_pos));
  else if (dtype == WF_BF16)
    WF_CHECK_CUDA(launch_pdl(3, embed_kernel<__nv_bfloat16>, grid, dim3(128), 0, stream, tokens, tok_stride, pos_ptr,
                             pos_const, tok_emb, pos_emb, (__nv_bfloat16*)out, ldo, d, n_pos));
  else
    WF_REQUIRE(false, "embed: bad dtype %d", dtype);
  count_launch();
  return WF_OK;
}

// ------------------------------------------------------------------ row-periodic table add (model.py:250, :322)
// out[r, :] = TO(float(in[r, :]) + table[r % mod, :])  - positional embedding added to features / stem output
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
add_rowmod_kernel(const TI* __restrict__ in, long long ldi, const float* __restrict__ table, TO* __restrict__ out,
                  long long ldo, long long rows, int d, int mod) {
  const long long total = rows * d;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / d;
    const int c = static_cast<int>(i % d);
    out[r * ldo + c] = from_f32<TO>(to_f32(in[r * ldi + c]) + table[(r % mod) * d + c]);
  }
}

int add_rowmod(int in_dtype, int out_dtype, const void* in, long long ldi, const float* table, void* out,
               long long ldo, long long rows, int d, int mod, cudaStream_t stream) {
  WF_REQUIRE(rows > 0 && d > 0 && mod > 0, "add_rowmod: bad shape");
  long long g = (rows * d + 255) / 256;
  const long long cap = 32LL * num_sms();
  if (g > cap) g = cap;
  const unsigned grid = static_cast<unsigned>(g);
  if (in_dtype == WF_F32 && out_dtype == WF_F32)
    add_rowmod_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)in, ldi, table, (float*)out, ldo, rows, d, mod);
  else if (in_dtype == WF_F32 && out_dtype == WF_BF16)
    add_rowmod_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)in, ldi, table, (__nv_bfloat16*)out, ldo, rows, d, mod);
  else if (in_dtype == WF_BF16 && out_dtype == WF_BF16)
    add_rowmod_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, ldi, table, (__nv_bfloat16*)out, ldo, rows, d, mod);
  else
    WF_REQUIRE(false, "add_rowmod: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

// ------------------------------------------------------------------ dtype casts (weight repack, feature ingest)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = from_f32<TO>(to_f32(in[i]));
}

int cast_copy(int in_dtype, int out_dtype, const void* in, void* out, long long n, cudaStream_t stream) {
  WF_REQUIRE(n > 0, "cast: empty input");
  long long g = (n + 255) / 256;
  const long long cap = 32LL * num_sms();
  if (g > cap) g = cap;
  const unsigned grid = static_cast<unsigned>(g);
  if (in_dtype == WF_F32 && out_dtype == WF_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)in, (__nv_bfloat16*)out, n);
  else if (in_dtype == WF_BF16 && out_dtype == WF_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, (float*)out, n);
  else
    WF_REQUIRE(false, "cast: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

}  // namespace wf
