// One-token cross-attention over the SOURCE rows themselves instead of over per-layer K/V caches ("absorbed"
// projections).  Replaces the per-step recompute of whisper/decoding.py:155-164 for the cross-attention
// (model.py:93-108 with xa) and the gated x-attention (model.py:110-134 with xt) of a greedy decode step.
//
// The reference computes, per layer and head h,  K_h = src Wk_h^T,  V_h = src Wv_h^T + bv_h  (src = encoder output
// xa [T, d] or the projected features xt [T_x, d]) and  o_h = softmax(q_h K_h^T / 8) V_h.  Caching K, V per layer
// costs 2 T d elements per (clip, layer) of HBM traffic at every step.  With
//     q'_h = Wk_h^T q_h   (d-vector)          scores_h = src q'_h / 8                 (same numbers as q_h K_h^T / 8)
//     c_h  = softmax(scores_h)^T src           o_h = Wv_h c_h + bv_h                   (sum of the weights is 1)
// the step streams src ONCE for all heads (T d elements per (clip, layer): half the bytes, and the same src serves
// every layer, so no per-layer K/V projection pass and no K/V arena exist at all); the extra arithmetic
// (2 x 2 T d H flop per clip) goes to the tensor cores.  Three kernels:
//   latent_query_kernel  q [R, d] -> q' [R, H, d]         (per head a [R,64] x [64,d] GEMM, mma.sync)
//   latent_attn_kernel   q', src  -> c  [R, H, d]         (tcgen05 + TMA + thread-block cluster, below)
//   latent_value_kernel  c -> o [R, d]                    (per head a [R,d] x [d,64] GEMM + bias, mma.sync)
//
// latent_attn_kernel: one CTA per clip, the whole latent width d in the CTA, 32-key tiles (a tile = 32 contiguous
// source rows = d / 128 ring stages of [32 keys x 128 columns], loaded by TMA exactly once, ring of ~1.9 tiles):
//   tcgen05  : S[64 rows (heads) x 32 keys] = q' (A, K-major, resident) x tile^T (B, K-major)        -> TMEM
//   softmax  : thread = head (its 32 scores in registers, no shuffles), lazy reference maximum, P^T row -> smem
//   tcgen05  : C^T[128 columns x 32 heads] += tile^T (A, MN-major: the SAME smem bytes) x P^T (B)    -> TMEM,
//              d / 128 accumulators; each stage is freed by the commit of the MMAs that read it
// and writes c_h = C^T[:, h] / l_h.  (A first version split d over a 5-CTA cluster and exchanged partial scores
// through DSMEM: 1.39 TB/s of source bytes, bound by the per-tile exchange chain; git 5357dfb.)
#include "common.cuh"
#include "kernels.h"

namespace wf {

static constexpr int LA_KT = 32;                    // keys per tile
static constexpr int LA_CHUNK = LA_KT * 128;        // 4 KB: [32 keys x 64 columns] bf16, 128B-swizzled = one TMA box
static constexpr int LA_STAGE = 2 * LA_CHUNK;       // 8 KB: 128 columns = the A operand of one context accumulator
static constexpr int LA_NH = 32;                    // head columns of the context MMAs (H <= 32)
static constexpr int LA_PT = LA_NH * 128;           // P^T operand: 32 rows (heads) x 128 B (64 keys; 32 used)
static constexpr int LA_MISC = 2048;                // alpha [2][32] | 1/l [32] | flags [2][2] | barriers
static constexpr int LA_MAX_STAGES = 28;
static constexpr int LA_SMEM_LIMIT = 227 * 1024;
static constexpr int LA_TMEM_COLS = 512;            // S (2 x 32) | C^T (d / 128 accumulators x 32)
static constexpr int LA_TMEM_C = 2 * LA_KT;
static constexpr int LA_THREADS = 256;

// non-blocking probe (mbarrier.try_wait may suspend the thread for a system-dependent time when the phase is still
// open, which a thread that polls two barriers cannot afford)
__device__ __forceinline__ bool la_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void la_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// bf16 x bf16 -> fp32, A MN-major (bit 15), B K-major
__host__ __device__ constexpr uint32_t la_idesc_a_mn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 15); }
// MN-major operand, 128B swizzle: rows = K index (128 B = 64 MN elements each), 8-row groups 1024 B apart (SBO),
// the next 64 MN elements one 4 KB chunk further (LBO)
__device__ __forceinline__ uint64_t la_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(LA_CHUNK >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(LA_THREADS, 1)
latent_attn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q,
                   __nv_bfloat16* __restrict__ ctx, int T, int H, int HP, int NS, int NST, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                              // NST stages of [32 keys x 128 columns]
  uint8_t* qs = ring + NST * LA_STAGE;               // q': d / 64 K-major atoms of HP rows (heads)
  const int q_atom = HP * 128;
  uint8_t* pt = qs + 2 * NS * q_atom;                // P^T operand, double-buffered
  uint8_t* misc = pt + 2 * LA_PT;
  float* al_buf = reinterpret_cast<float*>(misc);    // [2][32] rescale factor of each head for the tile
  float* linv_buf = al_buf + 64;                     // [32]
  int* flag_buf = reinterpret_cast<int*>(al_buf + 96);   // [2][2] "some head of this warp moved its reference maximum"
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 512);
  uint64_t* full = bars;                             // [NST] stage landed
  uint64_t* empty = bars + LA_MAX_STAGES;            // [NST] the context MMAs that read the stage completed
  uint64_t* q_full = bars + 2 * LA_MAX_STAGES;
  uint64_t* s_full = q_full + 1;                     // [2] scores of a tile in TMEM
  uint64_t* s_free = s_full + 2;                     // [2] ... copied to registers (64 arrivals)
  uint64_t* p_ready = s_free + 2;                    // [2] P^T staged, C^T rescaled (128 arrivals)
  uint64_t* c_done = p_ready + 2;                    // [2] context MMAs of a tile completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;                          // clip
  const int n_tiles = (T + LA_KT - 1) / LA_KT;
  const int d = NS * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_q);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 64); mbar_init(&p_ready[i], 128); mbar_init(&c_done[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<LA_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    const long long total = static_cast<long long>(n_tiles) * NS;
    int slot = 0, s = 0, j = 0;
    uint32_t phase = 0;
    bool q_sent = false;
    for (long long g = 0; g < total; ++g) {
      if (g == NST && !q_sent) {
        // the source rows are static, q' comes from the previous kernel: the first ring-full is requested before the
        // dependency wait
        pdl_wait();
        mbar_arrive_expect_tx(q_full, 2 * NS * q_atom);
        for (int c = 0; c < 2 * NS; ++c) tma_load_2d(qs + c * q_atom, &map_q, q_full, c * 64, b * H);
        q_sent = true;
      }
      mbar_wait(&empty[slot], phase ^ 1);
      mbar_arrive_expect_tx(&full[slot], LA_STAGE);
      tma_load_2d(ring + slot * LA_STAGE, &map_x, &full[slot], s * 128, b * T + j * LA_KT);
      tma_load_2d(ring + slot * LA_STAGE + LA_CHUNK, &map_x, &full[slot], s * 128 + 64, b * T + j * LA_KT);
      if (++s == NS) { s = 0; ++j; }
      if (++slot == NST) { slot = 0; phase ^= 1; }
    }
    if (!q_sent) {
      pdl_wait();
      mbar_arrive_expect_tx(q_full, 2 * NS * q_atom);
      for (int c = 0; c < 2 * NS; ++c) tma_load_2d(qs + c * q_atom, &map_q, q_full, c * 64, b * H);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    // scores  S[heads (64 rows, H valid) x 32 keys] = q' (A) x stage^T (B), both K-major, K = d in 16-column steps;
    // context C^T[128 columns x 32 heads] (+)= stage^T (A, MN-major: the same bytes) x P^T (B, K-major), K = 32 keys.
    // One stage of scores work is issued whenever its bytes have landed; a tile's context MMAs as soon as its P^T is
    // staged (they free the ring, so they go first).
#if defined(LA_EXP) && LA_EXP == 4
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, LA_KT);
#else
    constexpr uint32_t idesc_s = umma_idesc_bf16(64, LA_KT);
#endif
    constexpr uint32_t idesc_c = la_idesc_a_mn(128, LA_NH);
    const uint32_t ring_a = smem_u32(ring), qs_a = smem_u32(qs), pt_a = smem_u32(pt);
    int sj = 0, ss = 0, s_slot = 0, cj = 0, c_slot = 0;
    uint32_t s_phase = 0;
    mbar_wait(q_full, 0);
    long long t0 = clock64();
    while (cj < n_tiles) {
      bool progressed = false;
      if (la_test_wait(&p_ready[cj & 1], (cj >> 1) & 1)) {
        tc_fence_after();
        const uint32_t pb = pt_a + (cj & 1) * LA_PT;
        for (int a = 0; a < NS; ++a) {
          const uint32_t st = ring_a + c_slot * LA_STAGE;
#pragma unroll
          for (int kk = 0; kk < LA_KT / 16; ++kk)
#if defined(LA_EXP) && (LA_EXP == 1 || LA_EXP == 3)
            if (a < 0)
#endif
            umma_f16(tmem_base + LA_TMEM_C + a * LA_NH, la_desc_mn(st + kk * 2048), umma_desc_kmajor_sw128(pb) + 2 * kk,
                     idesc_c, (cj > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty[c_slot]);
          if (++c_slot == NST) c_slot = 0;
        }
        umma_commit(&c_done[cj & 1]);
        ++cj;
        progressed = true;
      }
      if (sj < n_tiles && la_test_wait(&full[s_slot], s_phase) &&
          (ss != 0 || sj < 2 || la_test_wait(&s_free[sj & 1], ((sj >> 1) - 1) & 1))) {
        tc_fence_after();
        const uint32_t st = ring_a + s_slot * LA_STAGE;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(qs_a + (2 * ss + c) * q_atom);
          const uint64_t b_desc = umma_desc_kmajor_sw128(st + c * LA_CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
#if defined(LA_EXP) && (LA_EXP == 2 || LA_EXP == 3)
            if ((c | k) == 0)
#endif
            umma_f16(tmem_base + (sj & 1) * LA_KT, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (ss | c | k) != 0);
        }
        if (++ss == NS) {
          umma_commit(&s_full[sj & 1]);
          ss = 0;
          ++sj;
        }
        if (++s_slot == NST) { s_slot = 0; s_phase ^= 1; }
        progressed = true;
      }
      if (progressed) {
        t0 = clock64();
      } else if (clock64() - t0 > 4000000000LL) {
        printf("libwf: latent attention MMA issuer timeout (block %d scores tile %d stage %d, context tile %d)\n",
               blockIdx.x, sj, ss, cj);
        __trap();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax (thread = head = TMEM lane of S)
    const int wq = warp - 4;
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    // a 64-row accumulator keeps rows 16 q .. 16 q + 15 in lanes 32 q .. 32 q + 15 (profiles/r01_probe_tmem_m64_layout)
    const int head = wq * 16 + lane;
    const bool act = wq < 2 && lane < 16 && head < H;
    float m_ref = -INFINITY, l_run = 0.f;
    pdl_wait();
    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      uint32_t pk[LA_KT / 2];
      if (wq < 2) {
        mbar_wait(&s_full[buf], ph);
        tc_fence_after();
        uint32_t sv[32];
        tmem_ld_32x32(lane_base + buf * LA_KT, sv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[buf]);
        const int nv = min(LA_KT, T - j * LA_KT);
        float mt = -INFINITY;
#pragma unroll
        for (int i = 0; i < LA_KT; ++i) {
          const float s = i < nv ? __uint_as_float(sv[i]) : -INFINITY;
          sv[i] = __float_as_uint(s);
          mt = fmaxf(mt, s);
        }
        // the reference maximum only moves when a tile exceeds it by more than 2^8 (p stays <= 256: exact enough in
        // bf16 / fp32), so the context accumulators are almost never rescaled
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_ref = mt;
        } else if ((mt - m_ref) * sl2 > 8.f) {
          alpha = ex2_approx((m_ref - mt) * sl2);
          m_ref = mt;
          need = act;
        }
        const float mb = m_ref * sl2;
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < LA_KT / 2; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -mb));
          pk[i] = pack_bf16(p0, p1);
          lsum += bf16lo(pk[i]) + bf16hi(pk[i]);      // the sums the tensor core will see
        }
        l_run = l_run * alpha + lsum;
        if (act) al_buf[buf * 32 + head] = alpha;
        const bool any = __any_sync(0xffffffffu, need);
        if (lane == 0) flag_buf[buf * 2 + wq] = any ? 1 : 0;
      }
      la_bar(1);
      const bool rescale = (flag_buf[buf * 2] | flag_buf[buf * 2 + 1]) != 0;
      if (j >= 2) mbar_wait(&c_done[buf], ((j - 2) >> 1) & 1);     // P^T[buf] is no longer read by tile j - 2
      if (rescale) {
        mbar_wait(&c_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        for (int a = 0; a < NS; ++a) {
          uint32_t cv[32];
          tmem_ld_32x32(lane_base + LA_TMEM_C + a * LA_NH, cv);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < LA_NH; ++h)
            if (h < H) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * al_buf[buf * 32 + h]);
          tmem_st_32x32(lane_base + LA_TMEM_C + a * LA_NH, cv);
        }
        tmem_st_wait();
      }
      if (act) {
        uint8_t* row = pt + buf * LA_PT + head * 128;
#pragma unroll
        for (int u = 0; u < LA_KT / 8; ++u)
          *reinterpret_cast<uint4*>(row + ((u ^ (head & 7)) << 4)) =
              make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
    }
    if (act) linv_buf[head] = 1.0f / l_run;
    la_bar(1);
    // ---- epilogue: ctx[b, h, 128 a + tid] = C^T[a][tid][h] / l_h   (thread = latent column)
    mbar_wait(&c_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    for (int a = 0; a < NS; ++a) {
      uint32_t cv[32];
      tmem_ld_32x32(lane_base + LA_TMEM_C + a * LA_NH, cv);
      tmem_ld_wait();
      __nv_bfloat16* out = ctx + static_cast<long long>(b) * H * d + a * 128 + tid;
#pragma unroll
      for (int h = 0; h < LA_NH; ++h)
        if (h < H) out[static_cast<long long>(h) * d] = __float2bfloat16_rn(__uint_as_float(cv[h]) * linv_buf[h]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<LA_TMEM_COLS>(tmem_base);
  }
}

int latent_attention(const void* qp, const void* src, void* ctx, int B, int T, int H, cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(B > 0 && T > 0 && H > 0 && H <= LA_NH && d % 128 == 0,
             "latent attention: needs head_dim 64, an even number of heads and at most 32 of them (got %d heads)", H);
  const int hp = (H + 7) / 8 * 8, ns = d / 128;
  const int fixed = 1024 + 2 * ns * hp * 128 + 2 * LA_PT + LA_MISC;
  int nst = (LA_SMEM_LIMIT - fixed) / LA_STAGE;
  if (nst > LA_MAX_STAGES) nst = LA_MAX_STAGES;
  WF_REQUIRE(nst > ns, "latent attention: the stage ring (%d) must hold more than one tile (%d stages)", nst, ns);
  const int smem = fixed + nst * LA_STAGE;
  CUtensorMap mx, mq;
  int rc = make_map_bf16(&mx, src, static_cast<long long>(B) * T, d, d, LA_KT);
  if (rc) return rc;
  rc = make_map_bf16(&mq, qp, static_cast<long long>(B) * H, d, d, hp);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LA_SMEM_LIMIT));
    configured = true;
  }
  const float sl2 = 0.125f * 1.44269504088896340736f;   // 64^-0.5 * log2(e)
  WF_CHECK_CUDA(launch_pdl(2, latent_attn_kernel, dim3(B), dim3(LA_THREADS), static_cast<size_t>(smem), stream, mx, mq,
                           reinterpret_cast<__nv_bfloat16*>(ctx), T, H, hp, ns, nst, sl2));
  count_launch();
  return WF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// q'[row, h, n] = sum_j q[row, 64 h + j] * wkT[n, 64 h + j]      (wkT = Wk^T, [d, d] row-major, packed once)
// grid (d / 128, H, ceil(R / 128)), 256 threads: warp w owns rows 16 w .. 16 w + 15 of the 128 x 128 output tile.
static constexpr int LQ_LD = 72;   // smem row stride in elements (144 B: conflict-free ldmatrix)

__global__ void __launch_bounds__(256)
latent_query_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ wkT,
                    __nv_bfloat16* __restrict__ qp, int R, int H, int d) {
  __shared__ __align__(16) __nv_bfloat16 As[128 * LQ_LD];
  __shared__ __align__(16) __nv_bfloat16 Bs[128 * LQ_LD];
  const int n0 = blockIdx.x * 128, h = blockIdx.y, m0 = blockIdx.z * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  for (int v = threadIdx.x; v < 128 * 8; v += 256) {
    const int row = v >> 3, c8 = (v & 7) * 8;
    uint4 a = make_uint4(0, 0, 0, 0);
    if (m0 + row < R) a = *reinterpret_cast<const uint4*>(q + (m0 + row) * ldq + h * 64 + c8);
    *reinterpret_cast<uint4*>(&As[row * LQ_LD + c8]) = a;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (n0 + row < d) w = *reinterpret_cast<const uint4*>(wkT + static_cast<long long>(n0 + row) * d + h * 64 + c8);
    *reinterpret_cast<uint4*>(&Bs[row * LQ_LD + c8]) = w;
  }
  __syncthreads();
  uint32_t af[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ldmatrix_x4(af[k], &As[(warp * 16 + (lane & 15)) * LQ_LD + k * 16 + (lane >> 4) * 8]);
  const int g = lane >> 2, t = lane & 3;
  const int row_lo = m0 + warp * 16 + g;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t bf[4];
      ldmatrix_x4(bf, &Bs[(nt * 8 + (lane & 7)) * LQ_LD + kp * 32 + (lane >> 3) * 8]);
      mma_bf16_16816(c, af[2 * kp], bf[0], bf[1]);
      mma_bf16_16816(c, af[2 * kp + 1], bf[2], bf[3]);
    }
    const int n = n0 + nt * 8 + 2 * t;
    if (n < d) {
      if (row_lo < R)
        *reinterpret_cast<uint32_t*>(qp + (static_cast<long long>(row_lo) * H + h) * d + n) = pack_bf16(c[0], c[1]);
      if (row_lo + 8 < R)
        *reinterpret_cast<uint32_t*>(qp + (static_cast<long long>(row_lo + 8) * H + h) * d + n) = pack_bf16(c[2], c[3]);
    }
  }
}

int latent_query(const void* q, long long ldq, const void* wkT, void* qp, int R, int H, cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(wkT) & 15) == 0,
             "latent query: operands must be 16-byte aligned");
  dim3 grid((d + 127) / 128, H, (R + 127) / 128);
  WF_CHECK_CUDA(launch_pdl(0, latent_query_kernel, grid, dim3(256), 0, stream,
                           reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(wkT),
                           reinterpret_cast<__nv_bfloat16*>(qp), R, H, d));
  count_launch();
  return WF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// o[row, 64 h + j] = sum_n c[row, h, n] * wv[64 h + j, n] + bv[64 h + j]
// grid (H, ceil(R / 16)), 128 threads; K = d streamed in 64-column chunks through a 4-stage cp.async ring;
// warp w owns output columns 16 w .. 16 w + 15 of the head.
static constexpr int LV_STAGES = 4;
static constexpr int LV_A = 16 * LQ_LD;      // elements per A stage
static constexpr int LV_B = 64 * LQ_LD;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t sz = pred ? 16u : 0u;     // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128)
latent_value_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ wv, long long ldw,
                    const float* __restrict__ bv, __nv_bfloat16* __restrict__ o, long long ldo, int R, int H, int d) {
  __shared__ __align__(16) __nv_bfloat16 As[LV_STAGES * LV_A];
  __shared__ __align__(16) __nv_bfloat16 Bs[LV_STAGES * LV_B];
  const int h = blockIdx.x, m0 = blockIdx.y * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = d / 64;
  const __nv_bfloat16* a_base = ctx + (static_cast<long long>(m0) * H + h) * d;       // row stride H * d
  const __nv_bfloat16* b_base = wv + static_cast<long long>(h) * 64 * ldw;
  auto load_b = [&](int kc, int st) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = threadIdx.x + i * 128;
      const int row = v >> 3, c8 = (v & 7) * 8;
      cp_async_16(&Bs[st * LV_B + row * LQ_LD + c8], b_base + row * ldw + kc * 64 + c8, true);
    }
  };
  auto load_a = [&](int kc, int st) {
    const int row = threadIdx.x >> 3, c8 = (threadIdx.x & 7) * 8;
    const bool ok = m0 + row < R;
    cp_async_16(&As[st * LV_A + row * LQ_LD + c8], a_base + (ok ? static_cast<long long>(row) * H * d : 0) + kc * 64 + c8,
                ok);
  };
  // the weights do not depend on the previous kernel: request them before the dependency wait
  for (int s = 0; s < LV_STAGES - 1; ++s)
    if (s < n_chunks) load_b(s, s);
  pdl_wait();
  for (int s = 0; s < LV_STAGES - 1; ++s) {
    if (s < n_chunks) load_a(s, s);
    cp_async_commit();     // group s = {A(s)} (+ all early B loads in group 0)
  }
  float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int kc = 0; kc < n_chunks; ++kc) {
    cp_async_wait<LV_STAGES - 2>();
    __syncthreads();
    const int nx = kc + LV_STAGES - 1;
    if (nx < n_chunks) { load_b(nx, nx % LV_STAGES); load_a(nx, nx % LV_STAGES); }
    cp_async_commit();
    const int st = kc % LV_STAGES;
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t a0[4], a1[4];
      ldmatrix_x4(a0, &As[st * LV_A + (lane & 15) * LQ_LD + kp * 32 + (lane >> 4) * 8]);
      ldmatrix_x4(a1, &As[st * LV_A + (lane & 15) * LQ_LD + kp * 32 + 16 + (lane >> 4) * 8]);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        uint32_t bf[4];
        ldmatrix_x4(bf, &Bs[st * LV_B + ((warp * 2 + nt) * 8 + (lane & 7)) * LQ_LD + kp * 32 + (lane >> 3) * 8]);
        mma_bf16_16816(c[nt], a0, bf[0], bf[1]);
        mma_bf16_16816(c[nt], a1, bf[2], bf[3]);
      }
    }
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = h * 64 + (warp * 2 + nt) * 8 + 2 * t;
    const float b0 = bv ? bv[col] : 0.f, b1 = bv ? bv[col + 1] : 0.f;
    if (m0 + g < R) *reinterpret_cast<uint32_t*>(o + (m0 + g) * ldo + col) = pack_bf16(c[nt][0] + b0, c[nt][1] + b1);
    if (m0 + g + 8 < R)
      *reinterpret_cast<uint32_t*>(o + (m0 + g + 8) * ldo + col) = pack_bf16(c[nt][2] + b0, c[nt][3] + b1);
  }
}

int latent_value(const void* ctx, const void* wv, long long ldw, const float* bv, void* o, long long ldo, int R, int H,
                 cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(ldw % 8 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(wv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 3) == 0,
             "latent value: operands must be 16-byte aligned");
  dim3 grid(H, (R + 15) / 16);
  WF_CHECK_CUDA(launch_pdl(0, latent_value_kernel, grid, dim3(128), 0, stream,
                           reinterpret_cast<const __nv_bfloat16*>(ctx), reinterpret_cast<const __nv_bfloat16*>(wv), ldw,
                           bv, reinterpret_cast<__nv_bfloat16*>(o), ldo, R, H, d));
  count_launch();
  return WF_OK;
}

}  // namespace wf
