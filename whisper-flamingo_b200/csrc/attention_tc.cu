// Flash attention on the 5th-gen tensor cores (bf16, head_dim 64, non-causal): the encoder's self-attention
// (T = 1500) and every long cross-attention of the teacher-forced decoder pass.
// Replaces reference whisper/model.py:93-108; scores live only in TMEM, never in HBM.
//
// One CTA per (batch, head, 256-query tile) = two 128-query warpgroups A and B sharing every K/V tile:
//   warp 0 lane 0 : TMA producer  - Q_A, Q_B once, then K/V tiles of 128 keys through a 3-stage ring
//   warp 1 lane 0 : MMA issuer    - S_w(j) = Q_w K(j)^T (128x128x64, fp32 in TMEM) as soon as S_w(j-1) has been read
//   warp 3 lane 0 : MMA issuer    - O_w += P_w(j) V(j) (A = P_w from tensor memory, B = V as an MN-major operand) as
//                                   soon as P_w(j) exists; the tensor pipe works for one warpgroup while the other does
//                                   its exponentials
//   warp 2        : TMEM allocator (S_A | S_B | O_A | O_B = 384 -> 512 columns)
//   warps 4..11   : softmax warpgroups: one thread owns one query row = one TMEM lane, so row max / row sum need
//                   no shuffles; exp2 with the 1/sqrt(64) scale folded in; P written as bf16 into a 128B-swizzled
//                   K-major tile; O rescaled in TMEM (tcgen05.ld / tcgen05.st) only when the running max of some
//                   row of the warp moved.
#include "common.cuh"
#include "kernels.h"

namespace wf {

// 2^x on the FMA / ALU pipes: round-to-nearest split x = n + f through the 1.5 * 2^23 trick, degree-3 polynomial for 2^f
// on [-0.5, 0.5] (max relative error 1.0e-4), n added to the exponent field.  x <= 126; anything below -126 gives 2^-126.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.055922035f, 0.24264008f);
  p = fmaf(p, f, 0.69312102f);
  p = fmaf(p, f, 0.99992448f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}


static constexpr int FT_M = 128;      // queries per softmax warpgroup
static constexpr int FT_WG = 2;       // warpgroups (query tiles) per CTA: 256 queries share every K/V tile
static constexpr int FT_N = 128;      // keys per tile
static constexpr int FT_D = 64;       // head dim
// P (the bf16 probabilities) goes from the softmax threads to the P V MMA through TENSOR memory (tcgen05.st, A operand
// read from TMEM) instead of a swizzled shared-memory tile: per K/V tile the CTA's shared memory then carries 96 KB
// (TMA fill + K and V operand reads of two warpgroups + Q) instead of 256 KB - at 128 B/clk the smem-P version spent as
// long on shared-memory traffic (2048 clk) as on its exponentials.  The freed 64 KB become two more K/V stages.
static constexpr bool FT_P_IN_TMEM = true;
#ifndef FT_MXC
#define FT_MXC 16     // independent FMNMX chains of the row maximum (4: 2262 us, 8: 2258, 16: 2236)
#endif
#ifndef FT_POLY
#define FT_POLY 0      // every FT_POLY-th pair of exponentials on the FMA / ALU pipes instead of MUFU.  Measured, B = 128
                       // x 20 heads x 1500: 0 (all MUFU) 2238 us = 659 TFLOP/s, 4: 2324 us, 3: 2432 us, 2: 2679 us - the
                       // softmax warps are bound by their issue slots, not by the MUFU pipe; MUFU stays the cheapest exp
#endif
// Measured and rejected (round 2): PERSISTENT CTAs (one per SM over the (query tile, head, clip) items, all barrier phases on
// a global tile counter, double-buffered Q, an o_free barrier so that producer and score issuer run into the next item
// during the epilogue): tokens identical, 2322 us (635 TFLOP/s) against 2200 us with one CTA per item; 2268 us with the
// warpgroup lag re-applied at every item, 2306 us with setmaxnreg (56 / 216 registers for the helper / softmax
// warpgroups).  The 2.5 us a fresh CTA waits for its first Q / K / V are evidently not the loss they look like.
// Also measured: probabilities in fp16 through ex2.approx.f16x2 - ptxas emits TWO scalar MUFU.EX2.F16 per f16x2 (no
// two-for-one on sm_100a), and tcgen05.mma kind::f16 with an fp16 A (P) and a bf16 B (V) is an illegal instruction.
#ifndef FT_PACK
#define FT_PACK 1      // packed f32x2 scale / row sums and three-input maxima in the softmax (it is issue bound)
#endif
#ifndef FT_LAZY
#define FT_LAZY 1      // lazy reference maximum (0: 2334 us, 1: 2267)
#endif
#ifndef FT_LAG
#define FT_LAG 400    // clk between the first score tiles of the two warpgroups (0: 2670 us, 400: 2332, 800: 2347, 1200: 2372)
#endif
static constexpr int FT_STAGES = FT_P_IN_TMEM ? 5 : 3;
static constexpr int FT_TILE = FT_N * FT_D * 2;             // 16 KB: one K (or V, or Q, or half-P) tile
static constexpr int FT_OFF_KV = FT_WG * FT_TILE;           // after Q_A, Q_B
static constexpr int FT_OFF_P = FT_OFF_KV + FT_STAGES * 2 * FT_TILE;
static constexpr int FT_OFF_BAR = FT_OFF_P + (FT_P_IN_TMEM ? 0 : FT_WG * 2 * FT_TILE);
static constexpr int FT_SMEM = FT_OFF_BAR + 256 + 1024;
static constexpr int FT_TMEM_COLS = 512;
static constexpr int FT_COL_O = 256;                         // S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384)
static constexpr int FT_COL_P = 384;                         // P_A [384,448) P_B [448,512): 128 keys x bf16 = 64 cells
static constexpr int FT_THREADS = 128 + FT_WG * 128;

__global__ void __launch_bounds__(FT_THREADS, 1)
fa_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
             const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ o, long long ldo, int Tq, int Tk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FT_OFF_BAR);
  uint64_t* q_full = bars;                           // [1]
  uint64_t* kv_full = bars + 1;                      // [STAGES]
  uint64_t* kv_empty = kv_full + FT_STAGES;          // [STAGES]
  uint64_t* s_full = kv_empty + FT_STAGES;           // [WG]  S_w(j) landed in TMEM
  uint64_t* s_free = s_full + FT_WG;                 // [WG]  S_w(j) has been read into registers
  uint64_t* p_full = s_free + FT_WG;                 // [WG]  P_w(j) in smem, O_w rescaled
  uint64_t* o_done = p_full + FT_WG;                 // [WG]  P_w(j) V(j) accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + FT_WG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (FT_WG * FT_M);
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (Tk + FT_N - 1) / FT_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FT_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < FT_WG; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128); mbar_init(&p_full[i], 128); mbar_init(&o_done[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<FT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_arrive_expect_tx(q_full, FT_WG * FT_TILE);
    for (int w = 0; w < FT_WG; ++w) tma_load_2d(smem + w * FT_TILE, &map_q, q_full, h * FT_D, b * Tq + q0 + w * FT_M);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % FT_STAGES;
      mbar_wait(&kv_empty[st], ((j / FT_STAGES) & 1) ^ 1);
      mbar_arrive_expect_tx(&kv_full[st], 2 * FT_TILE);
      uint8_t* dst = smem + FT_OFF_KV + st * 2 * FT_TILE;
      tma_load_2d(dst, &map_k, &kv_full[st], h * FT_D, b * Tk + j * FT_N);
      tma_load_2d(dst + FT_TILE, &map_v, &kv_full[st], h * FT_D, b * Tk + j * FT_N);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(FT_M, FT_N);        // S = Q K^T : both K-major
    auto issue_s = [&](int w, int j) {   // S_w(j) = Q_w K(j)^T  (kv_full(j) already observed by the caller)
      const int st = j % FT_STAGES;
      const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(smem + w * FT_TILE));
      const uint64_t k_desc = umma_desc_kmajor_sw128(smem_u32(smem + FT_OFF_KV + st * 2 * FT_TILE));
#pragma unroll
      for (int k = 0; k < FT_D / 16; ++k)
        umma_f16(tmem_base + w * FT_N, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k > 0);
      umma_commit(&s_full[w]);
    };
    // Two issuing threads: this one only computes scores, warp 3 only accumulates P V.  With one thread walking
    // S_A(j+1), PV_A(j), S_B(j+1), PV_B(j) in order, the scores of a warpgroup were issued only after the OTHER
    // warpgroup's probabilities had arrived: each warpgroup idled ~750-850 of its ~3700 clk per tile waiting for S.
    mbar_wait(q_full, 0);
#ifdef FT_TIMING
    long long ti_kv = 0, ti_sf = 0, ti_is = 0, tli = clock64();
#define FT_TI(acc_) do { const long long n_ = clock64(); acc_ += n_ - tli; tli = n_; } while (0)
#else
#define FT_TI(acc_)
#endif
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&kv_full[j % FT_STAGES], (j / FT_STAGES) & 1);
      FT_TI(ti_kv);
      for (int w = 0; w < FT_WG; ++w) {
        if (j > 0) mbar_wait(&s_free[w], (j - 1) & 1);   // the warpgroup has pulled S_w(j-1) out of TMEM
        FT_TI(ti_sf);
#if FT_LAG > 0
        if (j == 0 && w == 1) {
          // start warpgroup B half a period behind A: both spend half of a tile on the MUFU pipe, which they share
          mbar_wait(&s_free[0], 0);
          const long long t0 = clock64();
          while (clock64() - t0 < FT_LAG) {}
        }
#endif
        tc_fence_after();
        issue_s(w, j);
        FT_TI(ti_is);
      }
    }
#ifdef FT_TIMING
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 5)
      printf("fa S issuer per tile: wait K/V %lld | wait s_free %lld | issue %lld\n", ti_kv / n_tiles, ti_sf / n_tiles, ti_is / n_tiles);
#endif
  } else if (warp == 3 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer, O += P V
    constexpr uint32_t idesc_o = umma_idesc_bf16(FT_M, FT_D, 1);     // B (= V) MN-major
#ifdef FT_TIMING
    long long tpv_w = 0, tpv_i = 0;
#endif
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % FT_STAGES;
      const uint32_t v_addr = smem_u32(smem + FT_OFF_KV + st * 2 * FT_TILE + FT_TILE);
      for (int w = 0; w < FT_WG; ++w) {
#ifdef FT_TIMING
        const long long tp0 = clock64();
#endif
        mbar_wait(&p_full[w], j & 1);
#ifdef FT_TIMING
        const long long tp1 = clock64();
        tpv_w += tp1 - tp0;
#endif
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < FT_N / 16; ++kk) {
          // B: V rows 16 kk .. 16 kk + 15 (two 1024-B groups); A: 16 keys = 8 TMEM cells of this warpgroup's P region
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_addr + kk * 2048);
          if constexpr (FT_P_IN_TMEM) {
            umma_f16_ts(tmem_base + FT_COL_O + w * FT_D, tmem_base + FT_COL_P + w * (FT_N / 2) + kk * 8, b_desc, idesc_o,
                        (j > 0 || kk > 0) ? 1u : 0u);
          } else {
            const uint32_t p_addr = smem_u32(smem + FT_OFF_P + w * 2 * FT_TILE);
            const uint64_t a_desc = umma_desc_kmajor_sw128(p_addr + (kk >> 2) * FT_TILE) + 2 * (kk & 3);
            umma_f16(tmem_base + FT_COL_O + w * FT_D, a_desc, b_desc, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&o_done[w]);
#ifdef FT_TIMING
        tpv_i += clock64() - tp1;
#endif
      }
      umma_commit(&kv_empty[st]);  // both warpgroups are done with K(j), V(j) (their S(j) completed before P(j) existed)
    }
#ifdef FT_TIMING
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 5)
      printf("fa PV issuer per tile: wait P %lld | issue %lld\n", tpv_w / n_tiles, tpv_i / n_tiles);
#endif
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int w = (warp - 4) >> 2;                                   // warpgroup = query tile
    const int qd = warp & 3;
    const int r = qd * 32 + lane;                                   // query row in the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const uint32_t s_addr = lane_addr + w * FT_N;
    const uint32_t o_addr = lane_addr + FT_COL_O + w * FT_D;
    uint8_t* p_buf = smem + FT_OFF_P + (FT_P_IN_TMEM ? 0 : w * 2 * FT_TILE);
    (void)p_buf;
    const float sl2 = 0.125f * 1.44269504088896340736f;              // 64^-0.5 * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
#ifdef FT_TIMING
    long long t_w = 0, t_ld = 0, t_max = 0, t_exp = 0, t_od = 0, t_st = 0, tl = clock64();
    const long long tt0 = tl;
#define FT_T(acc_) do { const long long n_ = clock64(); acc_ += n_ - tl; tl = n_; } while (0)
#else
#define FT_T(acc_)
#endif
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int kbase = j * FT_N;
      const bool tail = kbase + FT_N > Tk;
      mbar_wait(&s_full[w], ph);
      FT_T(t_w);
      tc_fence_after();
      // one pass over TMEM: the whole 128-wide score row of this thread goes to registers
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(s_addr + c * 32, sv[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[w]);                                       // the tensor core may overwrite S_w now
      FT_T(t_ld);
      if (tail) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= Tk) sv[c][i] = 0xff800000u;  // -inf: masked key
      }
      // FT_MXC independent chains (one warp per SMSP: dependency latency is not hidden by TLP)
      float mxc[FT_MXC];
#pragma unroll
      for (int i = 0; i < FT_MXC; ++i) mxc[i] = -INFINITY;
      mxc[0] = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
#if FT_PACK
        for (int i = 0; i < 32; i += 2)
          mxc[(i >> 1) % FT_MXC] = fmax3(mxc[(i >> 1) % FT_MXC], __uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1]));
#else
        for (int i = 0; i < 32; ++i) mxc[i % FT_MXC] = fmaxf(mxc[i % FT_MXC], __uint_as_float(sv[c][i]));
#endif
      float mx = mxc[0];
#pragma unroll
      for (int i = 1; i < FT_MXC; ++i) mx = fmaxf(mx, mxc[i]);
#if FT_LAZY
      // the reference of a row only moves when its maximum leaves a window of 2^8 above it (probabilities stay below 256:
      // exact in the fp32 row sum, harmless in bf16): the O rescale - two TMEM round trips - disappears from most tiles
      if ((mx - m_run) * sl2 <= 8.0f) mx = m_run;                    // never on the first tile (m_run = -inf)
#endif
      const float alpha = ex2_approx((m_run - mx) * sl2);            // 0 on the first tile; 1 when the reference stays
      const float msc = mx * sl2;
      FT_T(t_max);
      // p = exp2(s * sl2 - m * sl2), packed to bf16 pairs in place (overlaps the P V MMA of the previous tile)
#if FT_PACK
      uint64_t ps2[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t sl2p = f2_pack(sl2, sl2), nmsc = f2_pack(-msc, -msc);
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sl2p, nmsc), x0, x1);
          const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
          ps2[(i >> 1) & 3] = f2_add(ps2[(i >> 1) & 3], f2_pack(e0, e1));
          sv[c][i >> 1] = pack_bf16(e0, e1);                          // slot i/2 <= i: already consumed
        }
      float psum;
      {
        float a, b;
        f2_unpack(f2_add(f2_add(ps2[0], ps2[1]), f2_add(ps2[2], ps2[3])), a, b);
        psum = a + b;
      }
#else
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = fmaf(__uint_as_float(sv[c][i]), sl2, -msc), x1 = fmaf(__uint_as_float(sv[c][i + 1]), sl2, -msc);
          // every FT_POLY-th pair takes the FMA / ALU pipes instead of MUFU (two softmax warps share a scheduler's
          // MUFU pipe: 8 clk per warp instruction; the polynomial is 8 instructions, 1e-4 relative - P is bf16)
          const bool poly = FT_POLY > 0 && ((i >> 1) % (FT_POLY > 0 ? FT_POLY : 1)) == (FT_POLY > 0 ? FT_POLY : 1) - 1;
          const float e0 = poly ? ex2_poly(x0) : ex2_approx(x0);
          const float e1 = poly ? ex2_poly(x1) : ex2_approx(x1);
          ps4[(i >> 1) & 3] += e0 + e1;
          sv[c][i >> 1] = pack_bf16(e0, e1);                          // slot i/2 <= i: already consumed
        }
      const float psum = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
#endif
      FT_T(t_exp);
      // O *= alpha (only when some row of this warp moved its max); PV_w(j-1) must have completed first
      if (j > 0) {
        mbar_wait(&o_done[w], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          uint32_t ov[2][32];
          tmem_ld_32x32(o_addr, ov[0]);
          tmem_ld_32x32(o_addr + 32, ov[1]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[c][i] = __float_as_uint(__uint_as_float(ov[c][i]) * alpha);
          tmem_st_32x32(o_addr, ov[0]);
          tmem_st_32x32(o_addr + 32, ov[1]);
          tmem_st_wait();
        }
      }
      l_run = l_run * alpha + psum;
      m_run = mx;
      FT_T(t_od);
      if constexpr (FT_P_IN_TMEM) {
        // P -> tensor memory: this thread's row, 64 cells of two bf16; the previous reader PV_w(j-1) has completed
        // (o_done was waited for above, or this is the first tile)
        uint32_t pk[2][32];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[c >> 1][(c & 1) * 16 + i] = sv[c][i];
        tmem_st_32x32(lane_addr + FT_COL_P + w * (FT_N / 2), pk[0]);
        tmem_st_32x32(lane_addr + FT_COL_P + w * (FT_N / 2) + 32, pk[1]);
        tmem_st_wait();
      } else {
        // P -> 128B-swizzled K-major tile (two 64-key halves); its previous reader PV_w(j-1) has completed
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* row_base = p_buf + (c >> 1) * FT_TILE + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ch = (c & 1) * 4 + g;                             // 16-byte chunk within the 128-byte row
            *reinterpret_cast<uint4*>(row_base + ((ch ^ (r & 7)) << 4)) =
                make_uint4(sv[c][g * 4], sv[c][g * 4 + 1], sv[c][g * 4 + 2], sv[c][g * 4 + 3]);
          }
        }
        fence_proxy_async_smem();   // P (generic-proxy stores) must be visible to the tensor core's async proxy
      }
      tc_fence_before();
      mbar_arrive(&p_full[w]);
      FT_T(t_st);
    }
#ifdef FT_TIMING
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 5 && lane == 0 && (warp & 3) == 0)
      printf("fa wg%d per tile: wait S %lld | tmem ld %lld | max %lld | exp %lld | wait PV(j-1) + rescale %lld | P store %lld ; total %lld\n",
             w, t_w / n_tiles, t_ld / n_tiles, t_max / n_tiles, t_exp / n_tiles, t_od / n_tiles, t_st / n_tiles,
             (clock64() - tt0) / n_tiles);
#endif
    // ---- epilogue: O / l -> bf16 -> global (one 128-byte row segment per thread)
    mbar_wait(&o_done[w], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int row = q0 + w * FT_M + r;
    __nv_bfloat16* orow = o + (static_cast<long long>(b) * Tq + row) * ldo + h * FT_D;
#pragma unroll 1
    for (int c = 0; c < FT_D / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(o_addr + c * 32, v);
      tmem_ld_wait();
      if (row < Tq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<FT_TMEM_COLS>(tmem_base);
  }
}

int attention_full_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      void* o, long long ldo, int B, int Tq, int Tk, int H, cudaStream_t stream) {
  WF_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention(tc): output must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc = make_map_bf16(&mq, q, static_cast<long long>(B) * Tq, static_cast<long long>(H) * FT_D, ldq, FT_M);
  if (rc) return rc;
  rc = make_map_bf16(&mk, k, static_cast<long long>(B) * Tk, static_cast<long long>(H) * FT_D, ldk, FT_N);
  if (rc) return rc;
  rc = make_map_bf16(&mv, v, static_cast<long long>(B) * Tk, static_cast<long long>(H) * FT_D, ldv, FT_N);
  if (rc) return rc;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(fa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
  }
  dim3 grid((Tq + FT_WG * FT_M - 1) / (FT_WG * FT_M), H, B);
  fa_tc_kernel<<<grid, FT_THREADS, FT_SMEM, stream>>>(mq, mk, mv, reinterpret_cast<__nv_bfloat16*>(o), ldo, Tq, Tk);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

}  // namespace wf
