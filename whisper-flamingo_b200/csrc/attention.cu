// Multi-head attention over full sequences (encoder self-attention T=1500, teacher-forced decoder passes).
// Replaces reference whisper/model.py:93-108 (qkv_attention): softmax_fp32((q s)(k s)^T + mask) v, s = 64^-0.25,
// without materialising the [B,H,Tq,Tk] fp32 score tensor the reference builds (and returns).
//
//   bf16: flash-style online softmax, S = Q K^T and O += P V on tensor cores (mma.sync m16n8k16, fp32 accumulate),
//         K/V tiles double-buffered in XOR-swizzled shared memory via cp.async.
//   fp32: CUDA-core kernel, one warp per query row, used by the token-exact fp32 engine.
// head_dim is fixed at 64 (every Whisper size).
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace wf {

static constexpr int HD = 64;

// ============================================================================ fp32 path
__global__ void __launch_bounds__(256)
attn_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                const float* __restrict__ v, long long ldv, float* __restrict__ o, long long ldo, int Tq, int Tk,
                int H, int causal) {
  const int warps_per_block = blockDim.x >> 5;
  const int tq = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (tq >= Tq) return;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const float scale = 0.125f;  // (64^-0.25)^2
  const float2 qv = *reinterpret_cast<const float2*>(q + (static_cast<long long>(b) * Tq + tq) * ldq + h * HD + lane * 2);
  const float* kb = k + static_cast<long long>(b) * Tk * ldk + h * HD + lane * 2;
  const float* vb = v + static_cast<long long>(b) * Tk * ldv + h * HD + lane * 2;
  const int kend = causal ? min(Tk, tq + 1) : Tk;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  int j = 0;
  for (; j + 4 <= kend; j += 4) {
    float s[4];
    float2 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 kk = *reinterpret_cast<const float2*>(kb + (j + u) * ldk);
      vv[u] = *reinterpret_cast<const float2*>(vb + (j + u) * ldv);
      s[u] = qv.x * kk.x + qv.y * kk.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = warp_sum(s[u]) * scale;
    const float mn = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
    const float corr = expf(m - mn);
    l *= corr; o0 *= corr; o1 *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float p = expf(s[u] - mn);
      l += p; o0 = fmaf(p, vv[u].x, o0); o1 = fmaf(p, vv[u].y, o1);
    }
    m = mn;
  }
  for (; j < kend; ++j) {
    const float2 kk = *reinterpret_cast<const float2*>(kb + j * ldk);
    const float2 vv = *reinterpret_cast<const float2*>(vb + j * ldv);
    const float s = warp_sum(qv.x * kk.x + qv.y * kk.y) * scale;
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), p = expf(s - mn);
    l = l * corr + p; o0 = o0 * corr + p * vv.x; o1 = o1 * corr + p * vv.y;
    m = mn;
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<float2*>(o + (static_cast<long long>(b) * Tq + tq) * ldo + h * HD + lane * 2) =
      make_float2(o0 * inv, o1 * inv);
}

// ============================================================================ bf16 path (mma.sync flash attention)
static constexpr int FA_BM = 64;   // queries per CTA (4 warps x 16 rows)
static constexpr int FA_BN = 64;   // keys per tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// tile of [rows][64] bf16, 128-byte rows, 16-byte chunks XOR-swizzled with (row & 7)
__device__ __forceinline__ __nv_bfloat16* sw_ptr(__nv_bfloat16* base, int row, int chunk) {
  return base + row * HD + ((chunk ^ (row & 7)) << 3);
}

// loads rows [r0, r0+64) x 64 columns of a [*, ld] matrix into a swizzled tile; rows >= r_end are zero-filled
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* tile, const __nv_bfloat16* g, long long ld, int r0,
                                                int r_end, int tid) {
#pragma unroll
  for (int i = 0; i < (64 * 8) / 128; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, chunk = idx & 7;
    const bool ok = (r0 + row) < r_end;
    const __nv_bfloat16* src = g + static_cast<long long>(ok ? (r0 + row) : 0) * ld + chunk * 8;
    cp_async16(sw_ptr(tile, row, chunk), src, ok);
  }
}

__global__ void __launch_bounds__(128)
attn_bf16_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                 long long ldk, const __nv_bfloat16* __restrict__ v, long long ldv, __nv_bfloat16* __restrict__ o,
                 long long ldo, int Tq, int Tk, int H, int causal) {
  __shared__ __align__(128) __nv_bfloat16 sQ[FA_BM * HD];
  __shared__ __align__(128) __nv_bfloat16 sK[2][FA_BN * HD];
  __shared__ __align__(128) __nv_bfloat16 sV[2][FA_BN * HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * FA_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qg = q + static_cast<long long>(b) * Tq * ldq + h * HD;
  const __nv_bfloat16* kg = k + static_cast<long long>(b) * Tk * ldk + h * HD;
  const __nv_bfloat16* vg = v + static_cast<long long>(b) * Tk * ldv + h * HD;

  int n_tiles = (Tk + FA_BN - 1) / FA_BN;
  if (causal) {
    const int last_q = min(q0 + FA_BM, Tq) - 1;
    n_tiles = min(n_tiles, last_q / FA_BN + 1);
  }

  load_tile_async(sQ, qg, ldq, q0, Tq, tid);
  load_tile_async(sK[0], kg, ldk, 0, Tk, tid);
  load_tile_async(sV[0], vg, ldv, 0, Tk, tid);
  cp_async_commit();

  // Q fragments: warp owns rows warp*16 .. +15; 4 k-steps of 16 over head_dim 64
  uint32_t qa[4][4];
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float row_m[2] = {-INFINITY, -INFINITY}, row_l[2] = {0.f, 0.f};
  const float sl2 = 0.125f * 1.44269504088896340736f;  // softmax scale * log2(e)

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_tile_async(sK[buf ^ 1], kg, ldk, (t + 1) * FA_BN, Tk, tid);
      load_tile_async(sV[buf ^ 1], vg, ldv, (t + 1) * FA_BN, Tk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(qa[ks], sw_ptr(sQ, row, chunk));
      }
    }
    // ---- S = Q K^T : 16 x 64 per warp = 8 n-tiles of 8 keys
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of n-tiles (16 keys)
        uint32_t kb[4];
        // matrices: (keys np*16+0..7, k chunk 2ks), (keys 0..7, chunk 2ks+1), (keys 8..15, chunk 2ks), (keys 8..15, chunk 2ks+1)
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kb, sw_ptr(sK[buf], row, chunk));
        mma_bf16_16816(s[np * 2], qa[ks], kb[0], kb[1]);
        mma_bf16_16816(s[np * 2 + 1], qa[ks], kb[2], kb[3]);
      }
    }
    // ---- mask (key tail, causal) + online softmax; thread owns rows r0 = lane/4 and r0 + 8
    const int kbase = t * FA_BN;
    const int qrow0 = q0 + warp * 16 + (lane >> 2);
    const bool need_mask = (kbase + FA_BN > Tk) || (causal && (kbase + FA_BN - 1 > q0 + warp * 16));
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = kbase + nt * 8 + (lane & 3) * 2 + (j & 1);
          const int qr = qrow0 + ((j >> 1) << 3);
          if (key >= Tk || (causal && key > qr)) s[nt][j] = -INFINITY;
        }
    }
    float mx[2] = {row_m[0], row_m[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_safe = (mx[r] == -INFINITY) ? 0.f : mx[r];  // fully masked row so far
      corr[r] = exp2f((row_m[r] - m_safe) * sl2);                // row_m = -inf -> 0
      msc[r] = m_safe * sl2;
      row_m[r] = mx[r];
      row_l[r] *= corr[r];
    }
    uint32_t pa[4][4];  // P as A fragments: 4 k-steps of 16 keys
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * sl2 - msc[0]);
      const float p1 = exp2f(s[nt][1] * sl2 - msc[0]);
      const float p2 = exp2f(s[nt][2] * sl2 - msc[1]);
      const float p3 = exp2f(s[nt][3] * sl2 - msc[1]);
      psum[0] += p0 + p1;
      psum[1] += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    row_l[0] += psum[0];
    row_l[1] += psum[1];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      oacc[dt][0] *= corr[0]; oacc[dt][1] *= corr[0];
      oacc[dt][2] *= corr[1]; oacc[dt][3] *= corr[1];
    }
    // ---- O += P V : k = keys (4 steps of 16), n = head_dim (8 tiles of 8); V needs the transposing ldmatrix
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of d-tiles (16 dims)
        uint32_t vb[4];
        // matrices: (keys ks*16+0..7, dims chunk 2dp), (keys 8..15, chunk 2dp), (keys 0..7, chunk 2dp+1), (keys 8..15, chunk 2dp+1)
        const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int chunk = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(vb, sw_ptr(sV[buf], row, chunk));
        mma_bf16_16816(oacc[dp * 2], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(oacc[dp * 2 + 1], pa[ks], vb[2], vb[3]);
      }
    }
    __syncthreads();  // everyone done with buf before the next prefetch overwrites it
  }

  // ---- finalize: row sums across the 4 lanes of a quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    row_l[r] += __shfl_xor_sync(0xffffffffu, row_l[r], 1);
    row_l[r] += __shfl_xor_sync(0xffffffffu, row_l[r], 2);
  }
  const float inv0 = row_l[0] > 0.f ? 1.0f / row_l[0] : 0.f;
  const float inv1 = row_l[1] > 0.f ? 1.0f / row_l[1] : 0.f;
  const int r0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* og = o + static_cast<long long>(b) * Tq * ldo + h * HD + (lane & 3) * 2;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    if (r0 < Tq)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r0) * ldo + dt * 8) =
          pack_bf16(oacc[dt][0] * inv0, oacc[dt][1] * inv0);
    if (r0 + 8 < Tq)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r0 + 8) * ldo + dt * 8) =
          pack_bf16(oacc[dt][2] * inv1, oacc[dt][3] * inv1);
  }
}

int attention_full(int dtype, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                   long long ldv, void* o, long long ldo, int B, int Tq, int Tk, int H, int causal,
                   cudaStream_t stream) {
  WF_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && H > 0, "attention: empty problem");
  WF_REQUIRE(B <= 65535 && H <= 65535, "attention: B/H too large for grid");
  if (dtype == WF_F32) {
    WF_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && ldo % 2 == 0, "attention(f32): strides must be even");
    dim3 grid((Tq + 7) / 8, H, B);
    attn_f32_kernel<<<grid, 256, 0, stream>>>((const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv,
                                              (float*)o, ldo, Tq, Tk, H, causal);
  } else if (dtype == WF_BF16) {
    WF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0,
               "attention(bf16): q/k/v row strides must be multiples of 8 elements");
    const auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    WF_REQUIRE(al(q) && al(k) && al(v), "attention(bf16): q/k/v must be 16-byte aligned");
    // long non-causal sequences (encoder self-attention, teacher-forced cross-attention): tcgen05 / TMEM kernel
    if (!causal && Tq >= 256 && Tk >= 256 && ldo % 8 == 0 && al(o) && getenv("WF_NO_TC_ATTENTION") == nullptr)
      return attention_full_tc(q, ldq, k, ldk, v, ldv, o, ldo, B, Tq, Tk, H, stream);
    dim3 grid((Tq + FA_BM - 1) / FA_BM, H, B);
    attn_bf16_kernel<<<grid, 128, 0, stream>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk,
                                               (const __nv_bfloat16*)v, ldv, (__nv_bfloat16*)o, ldo, Tq, Tk, H,
                                               causal);
  } else {
    WF_REQUIRE(false, "attention: bad dtype %d", dtype);
  }
  WF_CHECK_LAUNCH();
  return WF_OK;
}

}  // namespace wf
