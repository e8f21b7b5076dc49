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
//   latent_attn_kernel   q', src  -> c  [R, H, d]         (tcgen05 + TMA, below)
//   latent_value_kernel  c -> o [R, d]                    (per head a [R,d] x [d,64] GEMM + bias, mma.sync)
//
// latent_attn_kernel: one CTA per clip, the whole latent width d in the CTA, 128-key tiles, TWO passes over each tile.
// A tile is 128 keys x d = 320 KB at d = 1280: it cannot stay in shared memory between the scores and the context
// product.  The alternatives were built and measured first (git history): d split over a 5-CTA cluster with the partial
// scores exchanged through DSMEM (354 us at 128 clips x 1500 keys, large-v2), one pass over 32-key tiles with a
// 1.9-tile ring (320 us: the tile-synchronous release serialises load latency, softmax and ~4700 MMAs).  Here the
// second pass finds the tile in L2 - pass A is requested at most one tile ahead of pass B, first-pass loads carry an
// evict_last hint, second-pass loads evict_first - so HBM sees (nearly) every source row once.
//   pass A  TMA chunks [128 keys x 64 columns] + the matching 64-column atom of q' (ring A; q' is re-read from L2 for
//           every tile so that its 60 KB go to the rings) -> S^T[128 keys x 32 heads] += chunk (A, K-major) x q'^T (B)
//           with tcgen05.mma into TMEM; a slot is freed by the commit of its four MMAs
//   softmax thread = key; the reference maximum of a head only moves when a score exceeds it by more than 2^8, so the
//           common tile needs no cross-thread reduction at all; P^T (bf16) -> smem
//   pass B  TMA stages [128 keys x 128 columns] (ring B) -> C^T[128 columns x 32 heads] += stage^T (A, MN-major) x
//           P^T (B) with tcgen05.mma into TMEM, d / 128 accumulators
// Each pass has its own TMA thread and its own MMA-issuing thread.  What bounds it (profiles/r01_latent_*.txt): a
// narrow tcgen05.mma (N <= 64) occupies the SM's tensor front end for 40 clk whatever its shape and 53 clk when a
// single thread issues it (tools/probe/mma_cost.cu), so the 1920 MMAs of a clip cost >= 41 us; the wall is the data
// path: with q' resident the two rings held 144 KB, which at ~1.5 us of loaded HBM / L2 latency feeds ~100 GB/s per
// SM - the issuing threads waited ~400-500 clk per chunk (124 us; 99 us with all loads removed, 114 us with pass-A
// loads only).  Streaming q' with the chunks gives the rings 191 KB: 115 us.  L2 prefetch ahead of pass A and a
// longer lead of pass A both thrash L2 (133-178 us), the second pass must carry evict_first (152 us without), 64-key
// tiles pay more in hand-offs than they save in L2 misses (167 us), scores on mma.sync in the softmax warps cost
// ~32 clk per m16n8k16 (141 us).
// Output c_h = C^T[:, h] / l_h.
#include "common.cuh"
#include <cstdlib>
#include "kernels.h"

namespace wf {

#ifndef LA_KEYS_PER_TILE
#define LA_KEYS_PER_TILE 128
#endif
static constexpr int LA_KT = LA_KEYS_PER_TILE;      // keys per tile (128; 64 halves the L2 footprint, M = 64 score MMAs)
static_assert(LA_KT == 64 || LA_KT == 128, "score MMAs are 64 or 128 rows");
static constexpr int LA_CHUNK = LA_KT * 128;        // 16 KB: [128 keys x 64 columns] bf16, 128B-swizzled = one TMA box
static constexpr int LA_STAGE_B = 2 * LA_CHUNK;     // 32 KB: 128 columns = the A operand of one context accumulator
static constexpr int LA_NH = 32;                    // head columns of both MMAs (H <= 32)
static constexpr int LA_PATOM = LA_NH * 128;        // 4 KB: P^T rows (heads) x 64 keys
static constexpr int LA_PT = (LA_KT / 64) * LA_PATOM;   // P^T operand of a tile: 32 heads x LA_KT keys
static constexpr int LA_MISC = 3072;                // floats: m_ref[32] alpha[32] 1/l[32] red[4][32] | flags | barriers
static constexpr int LA_MAX_A = 16, LA_MAX_B = 6;
static constexpr int LA_SMEM_LIMIT = 227 * 1024;
#ifndef LA_ISSUERS
#define LA_ISSUERS 2       // MMA-issuing threads per pass.  A narrow tcgen05.mma occupies the tensor front end for 40 clk but
#endif                     // costs its issuing thread 53 clk, and a thread that waits for a chunk issues nothing: with two
                           // threads per pass (own accumulators) one waits while the other issues
static constexpr int LA_NI = LA_ISSUERS;
static_assert(LA_NI == 1 || LA_NI == 2, "one or two issuing threads per pass");
static constexpr int LA_TMEM_COLS = 512;            // S^T (2 buffers x LA_NI x 32) | C^T (d / 128 accumulators x 32)
static constexpr int LA_TMEM_C = 2 * LA_NI * LA_NH;
static constexpr int LA_THREADS = 256 + (LA_NI - 1) * 64;   // warps 8, 9: second issuer of pass A / pass B
#ifndef LA_PF
#define LA_PF 0        // chunks prefetched into L2 ahead of pass A (measured: 6 -> 133 us, 12 -> 141 us, 24 -> 178 us vs 124)
#endif
#ifndef LA_STREAM_Q
#define LA_STREAM_Q 1      // 1: the q' atom of a chunk travels with the chunk through ring A (re-read from L2 every tile)
#endif                     //    instead of all of q' staying resident: 60 KB more for the rings at d = 1280
                           //    (measured: resident, rings 5 + 2: 123 us; streamed, 7 + 2: 129 us, 5 + 3: 115 us, 4 + 4: 114 us)
#ifndef LA_AHEAD_PCT
#define LA_AHEAD_PCT 50    // how far beyond its own tile pass A may be requested ahead of pass B, % of a tile
#endif                     // (measured: 0 -> 190 us, 25 -> 128, 50 -> 123, 75 -> 129, 100 -> 136; more thrashes L2)


__device__ __forceinline__ void la_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ uint64_t la_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t la_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void la_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// bring one TMA box into L2 only: HBM requests in flight without a shared-memory slot behind each of them
__device__ __forceinline__ void la_tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}

// MN-major operand, 128B swizzle: rows = K index (128 B = 64 MN elements each), 8-row groups 1024 B apart (SBO), the
// next 64 MN elements one chunk further (LBO)
__device__ __forceinline__ uint64_t la_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(LA_CHUNK >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// bf16 x bf16 -> fp32, A MN-major (bit 15), B K-major
__host__ __device__ constexpr uint32_t la_idesc_a_mn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 15); }

__global__ void __launch_bounds__(LA_THREADS, 1)
latent_attn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q,
                   __nv_bfloat16* __restrict__ ctx, int T, int H, int HP, int NS, int NA, int NB, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_atom = HP * 128;
  const int slot_a = LA_CHUNK + (LA_STREAM_Q ? q_atom : 0);   // ring-A slot: chunk (+ its q' atom)
  uint8_t* ring_a = smem;                            // NA slots
  uint8_t* ring_b = ring_a + NA * slot_a;            // NB stages
  uint8_t* qs = ring_b + NB * LA_STAGE_B;            // q': d / 64 K-major atoms of HP rows (heads), unless streamed
  uint8_t* pt = qs + (LA_STREAM_Q ? 0 : 2 * NS * q_atom);   // P^T operand, double-buffered
  uint8_t* misc = pt + 2 * LA_PT;
  float* m_buf = reinterpret_cast<float*>(misc);     // [32] reference maximum of each head
  float* al_buf = m_buf + 32;                        // [32] rescale factor when the reference moved
  float* linv_buf = al_buf + 32;                     // [32]
  float* red = linv_buf + 32;                        // [4][32]
  int* flag_buf = reinterpret_cast<int*>(red + 128); // [2][4]
  volatile int* prog = flag_buf + 8;                 // [2] tiles fully requested by the pass-A / pass-B producer
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 1536);
  uint64_t* full_a = bars;                           // [NA]
  uint64_t* empty_a = full_a + LA_MAX_A;
  uint64_t* full_b = empty_a + LA_MAX_A;             // [NB]
  uint64_t* empty_b = full_b + LA_MAX_B;
  uint64_t* q_full = empty_b + LA_MAX_B;
  uint64_t* s_full = q_full + 1;                     // [2] scores of a tile in TMEM
  uint64_t* s_free = s_full + 2;                     // [2] ... copied to registers (128 arrivals)
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
    for (int i = 0; i < NA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    mbar_init(q_full, 1);
    prog[0] = 0;
    prog[1] = 0;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], LA_NI); mbar_init(&s_free[i], 128); mbar_init(&p_ready[i], 128); mbar_init(&c_done[i], LA_NI);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<LA_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer, pass A (chunks of 64 columns)
    // Pass B of tile j re-reads what pass A of tile j brought into L2, so the two request streams stay close: pass A
    // runs at most one tile plus LA_AHEAD_PCT % of a tile ahead of pass B, pass B never requests a tile before pass A
    // has (prog[0] / prog[1] = 64-column chunks requested by pass A / pass B).  One thread per ring.
#if defined(LA_HINT) && (LA_HINT & 1)
    const uint64_t keep = la_policy_evict_normal();
#else
    const uint64_t keep = la_policy_evict_last();
#endif
    int slot = 0, issued = 0;
    uint32_t phase = 0;
    const int q_after = n_tiles * 2 * NS < NA ? n_tiles * 2 * NS : NA;
    bool q_sent = false;
    auto send_q = [&]() {
      pdl_wait();       // q' comes from the previous kernel; the source rows are static
#if LA_STREAM_Q
      // the chunks requested so far wait for their q' atoms (slot i holds chunk i: issued <= NA)
      for (int i = 0; i < issued; ++i)
        tma_load_2d(ring_a + i * slot_a + LA_CHUNK, &map_q, &full_a[i], (i % (2 * NS)) * 64, b * H);
#else
      mbar_arrive_expect_tx(q_full, 2 * NS * q_atom);
      for (int i = 0; i < 2 * NS; ++i) tma_load_2d(qs + i * q_atom, &map_q, q_full, i * 64, b * H);
#endif
      q_sent = true;
    };
#if LA_PF > 0
    // L2 prefetch LA_PF chunks ahead of the ring: HBM requests in flight without a shared-memory slot behind them
    const int total = n_tiles * 2 * NS;
    int pj = 0, pc = 0, pg = 0;
    auto prefetch_next = [&]() {
      if (pg < total) {
        la_tma_prefetch_2d(&map_x, pc * 64, b * T + pj * LA_KT);
        ++pg;
        if (++pc == 2 * NS) { pc = 0; ++pj; }
      }
    };
    for (int i = 0; i < LA_PF; ++i) prefetch_next();
#endif
    const int lim = 2 * NS + (2 * NS * LA_AHEAD_PCT) / 100;
    for (int j = 0; j < n_tiles; ++j) {
      for (int c = 0; c < 2 * NS; ++c) {
        if (!q_sent && issued == q_after) send_q();
        if (issued - prog[1] >= lim) {
          const long long t0 = clock64();
          while (issued - prog[1] >= lim) {
            if (clock64() - t0 > 4000000000LL) {
              printf("libwf: latent attention pass-A producer timeout (block %d tile %d)\n", blockIdx.x, j);
              __trap();
            }
          }
        }
        mbar_wait(&empty_a[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_a[slot], slot_a);
        la_tma_load_2d(ring_a + slot * slot_a, &map_x, &full_a[slot], c * 64, b * T + j * LA_KT, keep);
#if LA_STREAM_Q
        if (q_sent) tma_load_2d(ring_a + slot * slot_a + LA_CHUNK, &map_q, &full_a[slot], c * 64, b * H);
#endif
#if LA_PF > 0
        prefetch_next();
#endif
        ++issued;
        prog[0] = issued;
        if (++slot == NA) { slot = 0; phase ^= 1; }
      }
    }
    if (!q_sent) send_q();
  } else if (warp == 2 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer, pass B (stages of 128 columns)
#if defined(LA_HINT) && (LA_HINT & 2)
    const uint64_t drop = la_policy_evict_normal();
#else
    const uint64_t drop = l2_policy_evict_first();
#endif
    int slot = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const long long t0 = clock64();
      while (prog[0] < (j + 1) * 2 * NS) {
        if (clock64() - t0 > 4000000000LL) {
          printf("libwf: latent attention pass-B producer timeout (block %d tile %d)\n", blockIdx.x, j);
          __trap();
        }
      }
      for (int a = 0; a < NS; ++a) {
        mbar_wait(&empty_b[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_b[slot], LA_STAGE_B);
        uint8_t* dst = ring_b + slot * LA_STAGE_B;
        la_tma_load_2d(dst, &map_x, &full_b[slot], a * 128, b * T + j * LA_KT, drop);
        la_tma_load_2d(dst + LA_CHUNK, &map_x, &full_b[slot], a * 128 + 64, b * T + j * LA_KT, drop);
        if (++slot == NB) { slot = 0; phase ^= 1; }
        prog[1] = (j * NS + a + 1) * 2;
      }
    }
  } else if ((warp == 1 || (LA_NI == 2 && warp == 8)) && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer(s), pass A
    // S^T[128 keys x 32 heads] = chunk (A, K-major) x q'^T (B, K-major; rows >= HP of an atom alias the next atom and
    // only produce head columns nobody reads), K = d in 16-column steps.  Issuer `me` takes the chunks with index
    // = me (mod LA_NI) into its own accumulator (the softmax adds them).
    constexpr uint32_t idesc_s = umma_idesc_bf16(LA_KT, LA_NH);
    const uint32_t ra = smem_u32(ring_a), qa = smem_u32(qs);
    const int me = warp == 1 ? 0 : 1;
#if !LA_STREAM_Q
    mbar_wait(q_full, 0);
#endif
#ifdef LA_TIMING
    long long tw = 0, tm = 0, tcm = 0, ts = 0, tt0 = clock64();
#define LA_T(acc) do { const long long n_ = clock64(); acc += n_ - tl; tl = n_; } while (0)
#else
#define LA_T(acc)
#endif
    for (int j = 0; j < n_tiles; ++j) {
#ifdef LA_TIMING
      long long tl = clock64();
#endif
      if (j >= 2) mbar_wait(&s_free[j & 1], ((j >> 1) - 1) & 1);
      LA_T(ts);
      const uint32_t acc = tmem_base + ((j & 1) * LA_NI + me) * LA_NH;
      for (int c = me; c < 2 * NS; c += LA_NI) {
        const int g = j * 2 * NS + c;          // chunk counter of the ring
        const int slot = g % NA;
        mbar_wait(&full_a[slot], (g / NA) & 1);
        LA_T(tw);
        tc_fence_after();
        const uint64_t a_desc = umma_desc_kmajor_sw128(ra + slot * slot_a);
        const uint64_t b_desc = umma_desc_kmajor_sw128(LA_STREAM_Q ? ra + slot * slot_a + LA_CHUNK : qa + c * q_atom);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(acc, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (c >= LA_NI || k > 0) ? 1u : 0u);
        LA_T(tm);
        umma_commit(&empty_a[slot]);
        LA_T(tcm);
      }
      umma_commit(&s_full[j & 1]);
    }
#ifdef LA_TIMING
    if (blockIdx.x == 0)
      printf("issuer A%d per chunk: wait full %lld | 4 MMAs %lld | commit %lld ; per tile: wait s_free %lld ; total/tile %lld\n",
             me, tw / (n_tiles * 2 * NS / LA_NI), tm / (n_tiles * 2 * NS / LA_NI), tcm / (n_tiles * 2 * NS / LA_NI),
             ts / n_tiles, (clock64() - tt0) / n_tiles);
#endif
  } else if ((warp == 3 || (LA_NI == 2 && warp == 9)) && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer(s), pass B
    // C^T[128 columns x 32 heads] (+)= stage^T (A, MN-major: rows = keys, two 64-column atoms 16 KB apart) x P^T (B,
    // K-major, two atoms of 64 keys), K = 128 keys in 16-key steps.  Issuer `me` takes the stages with ring counter
    // = me (mod LA_NI); every stage has its own accumulator.
    constexpr uint32_t idesc_c = la_idesc_a_mn(128, LA_NH);
    const uint32_t rb = smem_u32(ring_b), pa = smem_u32(pt);
    const int me = warp == 3 ? 0 : 1;
#ifdef LA_TIMING
    long long tw = 0, tm = 0, tcm = 0, ts = 0, tt0 = clock64();
#endif
    for (int j = 0; j < n_tiles; ++j) {
#ifdef LA_TIMING
      long long tl = clock64();
#endif
      mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
      LA_T(ts);
      tc_fence_after();
      const uint32_t pb = pa + (j & 1) * LA_PT;
      for (int a = 0; a < NS; ++a) {
        const int g = j * NS + a;              // stage counter of the ring
        if (LA_NI == 2 && (g & 1) != me) continue;
        const int slot = g % NB;
        mbar_wait(&full_b[slot], (g / NB) & 1);
        LA_T(tw);
        tc_fence_after();
        const uint32_t st = rb + slot * LA_STAGE_B;
#pragma unroll
        for (int kk = 0; kk < LA_KT / 16; ++kk)
          umma_f16(tmem_base + LA_TMEM_C + a * LA_NH, la_desc_mn(st + kk * 2048),
                   umma_desc_kmajor_sw128(pb + (kk >> 2) * LA_PATOM) + 2 * (kk & 3), idesc_c, (j > 0 || kk > 0) ? 1u : 0u);
        LA_T(tm);
        umma_commit(&empty_b[slot]);
        LA_T(tcm);
      }
      umma_commit(&c_done[j & 1]);
    }
#ifdef LA_TIMING
    if (blockIdx.x == 0)
      printf("issuer B%d per stage: wait full %lld | 8 MMAs %lld | commit %lld ; per tile: wait p_ready %lld ; total/tile %lld\n",
             me, tw / (n_tiles * NS / LA_NI), tm / (n_tiles * NS / LA_NI), tcm / (n_tiles * NS / LA_NI), ts / n_tiles,
             (clock64() - tt0) / n_tiles);
#endif
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ softmax (thread = key = TMEM lane of S^T)
    const int wq = warp - 4;
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float l_part[LA_NH];
#pragma unroll
    for (int h = 0; h < LA_NH; ++h) l_part[h] = 0.f;
    if (tid < LA_NH) m_buf[tid] = -INFINITY;
    pdl_wait();
    la_bar(1);
    // P^T element (head h, my key): atom = 64-key half, row = head (128 B), 16-byte units swizzled by the row
    // my key inside the tile: a 128-row accumulator keeps row r in lane r, a 64-row one keeps rows 16 q .. 16 q + 15
    // in lanes 32 q .. 32 q + 15 (profiles/r01_probe_tmem_m64_layout.txt)
    const int key = LA_KT == 128 ? tid : wq * 16 + (lane & 15);
    const bool lane_on = LA_KT == 128 || lane < 16;
    const uint32_t p_off = (key >> 6) * LA_PATOM + (key & 7) * 2;
    const uint32_t p_unit = (key & 63) >> 3;
    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[buf], ph);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(lane_base + buf * LA_NI * LA_NH, sv);
      if (LA_NI == 2) {    // the two issuers of pass A dealt the chunks to two accumulators
        uint32_t sv2[32];
        tmem_ld_32x32(lane_base + (buf * LA_NI + 1) * LA_NH, sv2);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < LA_NH; ++h) sv[h] = __float_as_uint(__uint_as_float(sv[h]) + __uint_as_float(sv2[h]));
      } else {
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_free[buf]);
      const bool valid = lane_on && j * LA_KT + key < T;
      // does any score leave the window of its head's reference maximum?  (always on the first tile)
      bool exceed = false;
#pragma unroll
      for (int h = 0; h < LA_NH; ++h)
        if (h < H) exceed = exceed || (__uint_as_float(sv[h]) - m_buf[h]) * sl2 > 8.f;
      const bool w_any = __any_sync(0xffffffffu, exceed && valid);
      if (lane == 0) flag_buf[buf * 4 + wq] = w_any ? 1 : 0;
      la_bar(1);
      const bool update = (flag_buf[buf * 4] | flag_buf[buf * 4 + 1] | flag_buf[buf * 4 + 2] | flag_buf[buf * 4 + 3]) != 0;
      if (update) {
        // move the references to the running maxima, rescale the sums and the context accumulators
#pragma unroll
        for (int h = 0; h < LA_NH; ++h)
          if (h < H) {
            const float mt = warp_max(valid ? __uint_as_float(sv[h]) : -INFINITY);
            if (lane == 0) red[wq * 32 + h] = mt;
          }
        la_bar(2);
        if (tid < H) {
          const float mt = fmaxf(fmaxf(red[tid], red[32 + tid]), fmaxf(red[64 + tid], red[96 + tid]));
          const float m_old = m_buf[tid];
          const float m_new = fmaxf(m_old, mt);
          al_buf[tid] = ex2_approx((m_old - m_new) * sl2);     // 0 on the first tile
          m_buf[tid] = m_new;
        }
        la_bar(3);
#pragma unroll
        for (int h = 0; h < LA_NH; ++h)
          if (h < H) l_part[h] *= al_buf[h];
        if (j > 0) {
          mbar_wait(&c_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          for (int a = 0; a < NS; ++a) {
            uint32_t cv[32];
            tmem_ld_32x32(lane_base + LA_TMEM_C + a * LA_NH, cv);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < LA_NH; ++h)
              if (h < H) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * al_buf[h]);
            tmem_st_32x32(lane_base + LA_TMEM_C + a * LA_NH, cv);
          }
          tmem_st_wait();
        }
      }
      if (j >= 2) mbar_wait(&c_done[buf], ((j - 2) >> 1) & 1);     // P^T[buf] is no longer read by tile j - 2
      uint8_t* prow = pt + buf * LA_PT + p_off;
#pragma unroll
      for (int h = 0; h < LA_NH; ++h)
        if (h < H) {
          const float p = valid ? ex2_approx((__uint_as_float(sv[h]) - m_buf[h]) * sl2) : 0.f;
          const __nv_bfloat16 pb = __float2bfloat16_rn(p);
          l_part[h] += __bfloat162float(pb);                       // the sums the tensor core will see
          *reinterpret_cast<__nv_bfloat16*>(prow + h * 128 + ((p_unit ^ (h & 7)) << 4)) = pb;
        }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
      // m_buf / al_buf / flag_buf[buf] are rewritten two tiles later at the earliest, behind la_bar(1) of tile j + 1
    }
    // ---- 1 / l
#pragma unroll
    for (int h = 0; h < LA_NH; ++h)
      if (h < H) {
        const float v = warp_sum(l_part[h]);
        if (lane == 0) red[wq * 32 + h] = v;
      }
    la_bar(2);
    if (tid < H) linv_buf[tid] = 1.0f / ((red[tid] + red[32 + tid]) + (red[64 + tid] + red[96 + tid]));
    la_bar(3);
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

int latent_attention(const void* qp, const void* src, void* ctx, long long part_stride, float* ml, int B, int T, int H,
                     cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(B > 0 && T > 0 && H > 0 && H <= LA_NH && d % 128 == 0 && LA_TMEM_C + (d / 128) * LA_NH <= LA_TMEM_COLS,
             "latent attention: needs head_dim 64, an even number of heads and at most %d of them (got %d heads)",
             2 * (LA_TMEM_COLS - LA_TMEM_C) / LA_NH, H);
  // default: the one-pass pair kernel (latent_pair.cu: 2-CTA clusters, tiles resident in shared memory);
  // WF_LATENT_PASS=2 keeps this file's two-pass kernel (A/B measurements; also the fallback when d % 256 != 0)
  static int passes = -1;
  if (passes < 0) {
    const char* e = getenv("WF_LATENT_PASS");
    passes = e ? atoi(e) : 1;
  }
  if (passes != 2) {
    const int rc1 = latent_attention_pair(qp, src, ctx, ml, part_stride, B, T, H, stream);
    if (rc1 != WF_ERR_UNSUPPORTED) return rc1;
  }
  WF_REQUIRE(ml == nullptr, "latent attention: the split form needs the pair kernel (wf_latent_split_supported)");
  const int hp = (H + 7) / 8 * 8, ns = d / 128;
  const int q_atom = hp * 128;
  const int fixed = 1024 + (LA_STREAM_Q ? 0 : 2 * ns * q_atom) + 2 * LA_PT + LA_MISC;
  const int slot_a = LA_CHUNK + (LA_STREAM_Q ? q_atom : 0);
  const int n = (LA_SMEM_LIMIT - fixed) / LA_CHUNK;       // chunk-sized units left for the two rings
#ifdef LA_NB_FORCE
  int nb = LA_NB_FORCE;
#else
  int nb = (LA_KT == 128 ? (n >= 12 ? (LA_NI == 2 ? 4 : 3) : 2) : 4);
#endif
  if (nb > LA_MAX_B) nb = LA_MAX_B;
  int na = (LA_SMEM_LIMIT - fixed - nb * LA_STAGE_B) / slot_a;
  if (na > LA_MAX_A) na = LA_MAX_A;
  // Two issuers per pass deal the ring slots by parity: a slot must always be consumed by the same thread (an mbarrier
  // wait is only valid at most one phase ahead, and TMA boxes complete out of order), so both rings are even.
  if (LA_NI == 2) { na &= ~1; nb &= ~1; }
  WF_REQUIRE(na >= 2, "latent attention: shared memory too small for the rings (%d chunks)", n);
  const int smem = fixed + na * slot_a + nb * LA_STAGE_B;
  CUtensorMap mx, mq;
  int rc = make_map_bf16(&mx, src, static_cast<long long>(B) * T, d, d, LA_KT);
  if (rc) return rc;
  rc = make_map_bf16(&mq, qp, static_cast<long long>(B) * H, d, d, hp);
  if (rc) return rc;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LA_SMEM_LIMIT));
  }
  const float sl2 = 0.125f * 1.44269504088896340736f;   // 64^-0.5 * log2(e)
  WF_CHECK_CUDA(launch_pdl(2, latent_attn_kernel, dim3(B), dim3(LA_THREADS), static_cast<size_t>(smem), stream, mx, mq,
                           reinterpret_cast<__nv_bfloat16*>(ctx), T, H, hp, ns, na, nb, sl2));
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
  // (no early launch of the dependents here: measured, the 148 attention CTAs then take the SMs the second wave of
  // this kernel's 200 CTAs needs and the pair costs 6 us more)
  // the weight tile does not depend on the previous kernel: its HBM round trip runs under the dependency wait
  uint4 w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = threadIdx.x + i * 256;
    const int row = v >> 3, c8 = (v & 7) * 8;
    w[i] = make_uint4(0, 0, 0, 0);
    if (n0 + row < d) w[i] = *reinterpret_cast<const uint4*>(wkT + static_cast<long long>(n0 + row) * d + h * 64 + c8);
  }
  pdl_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = threadIdx.x + i * 256;
    const int row = v >> 3, c8 = (v & 7) * 8;
    uint4 a = make_uint4(0, 0, 0, 0);
    if (m0 + row < R) a = *reinterpret_cast<const uint4*>(q + (m0 + row) * ldq + h * 64 + c8);
    *reinterpret_cast<uint4*>(&As[row * LQ_LD + c8]) = a;
    *reinterpret_cast<uint4*>(&Bs[row * LQ_LD + c8]) = w[i];
  }
  __syncthreads();
  uint32_t af[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ldmatrix_x4(af[k], &As[(warp * 16 + (lane & 15)) * LQ_LD + k * 16 + (lane >> 4) * 8]);
  const int g = lane >> 2, t = lane & 3;
  // The 16 x 128 outputs of a warp leave in two halves of 64 columns, staged in the warp's own rows of As (its A
  // fragments are in registers by now) so that every store instruction writes four full 128-byte row segments
  // (fragment layout: 16-byte pieces of eight different rows per instruction).
  __nv_bfloat16* stage = &As[(warp * 16) * LQ_LD];
  __syncwarp();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int nq = 0; nq < 8; ++nq) {
      const int nt = half * 8 + nq;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t bf[4];
        ldmatrix_x4(bf, &Bs[(nt * 8 + (lane & 7)) * LQ_LD + kp * 32 + (lane >> 3) * 8]);
        mma_bf16_16816(c, af[2 * kp], bf[0], bf[1]);
        mma_bf16_16816(c, af[2 * kp + 1], bf[2], bf[3]);
      }
      *reinterpret_cast<uint32_t*>(&stage[g * LQ_LD + nq * 8 + 2 * t]) = pack_bf16(c[0], c[1]);
      *reinterpret_cast<uint32_t*>(&stage[(g + 8) * LQ_LD + nq * 8 + 2 * t]) = pack_bf16(c[2], c[3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 4 + (lane >> 3), ch = lane & 7;         // row of the warp's 16, 16-byte chunk of the 128-byte half row
      const int row = m0 + warp * 16 + r, n = n0 + half * 64 + ch * 8;
      if (row < R && n < d)
        *reinterpret_cast<uint4*>(qp + (static_cast<long long>(row) * H + h) * d + n) =
            *reinterpret_cast<const uint4*>(&stage[r * LQ_LD + ch * 8]);
    }
    __syncwarp();
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
// grid (H, ceil(R / 16)), 128 threads; K = d streamed in 128-column stages through a 3-stage cp.async ring;
// warp w owns output columns 16 w .. 16 w + 15 of the head.
// K = d streamed in 128-column stages (64-column stages: twice the pipeline rounds, each a cp.async wait + block
// barrier - measured 8.2 / 11.8 us per launch for the plain / split form at large-v2 width)
static constexpr int LV_STAGES = 3;
static constexpr int LV_KC = 128;            // columns per stage
static constexpr int LV_LD = LV_KC + 8;      // smem row stride in elements (272 B: conflict-free ldmatrix)
static constexpr int LV_A16 = 16 * LV_LD;     // elements per 16-row A tile and stage
static constexpr int LV_B = 64 * LV_LD;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t sz = pred ? 16u : 0u;     // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// SPLIT: ctx holds two partial contexts per (row, head) `part_stride` elements apart, each normalised by its own row
// sum, with ml[(part * R + row) * 32 + head] = (reference maximum in log2 units, row sum; 0 = part absent) - what the
// persistent pair kernel (latent_pair.cu) leaves when a clip is cut at a cluster border.  The projection is linear in
// the context: both parts are projected and blended with the softmax weights of their segments.
// LV_MT = 16-row MMA tiles per CTA: 16 rows while H x ceil(R / 16) CTAs fit the SMs at once, else 32 (128 rows x 20 heads
// = 160 CTAs of 16 rows would put two CTAs on 12 of the SMs: 8.5 -> 7.3 us)
template <bool SPLIT, int LV_MT>
__global__ void __launch_bounds__(128)
latent_value_kernel(const __nv_bfloat16* __restrict__ ctx, long long part_stride, const float2* __restrict__ ml,
                    const __nv_bfloat16* __restrict__ wv, long long ldw, const float* __restrict__ bv,
                    __nv_bfloat16* __restrict__ o, long long ldo, int R, int H, int d) {
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int NST = LV_STAGES;
  constexpr int LV_ROWS = 16 * LV_MT, LV_A = LV_MT * LV_A16;
  extern __shared__ __align__(16) uint8_t lv_smem[];       // NST x (NP A tiles + one B tile): 65 KB, 78 KB in the split form
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(lv_smem);
  __nv_bfloat16* Bs = As + NP * NST * LV_A;
  __shared__ float wgt[2][LV_ROWS];
  const int h = blockIdx.x, m0 = blockIdx.y * LV_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = (d + LV_KC - 1) / LV_KC;
  const __nv_bfloat16* a_base = ctx + (static_cast<long long>(m0) * H + h) * d;       // row stride H * d
  const __nv_bfloat16* b_base = wv + static_cast<long long>(h) * 64 * ldw;
  auto load_b = [&](int kc, int st) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = threadIdx.x + i * 128;
      const int row = v >> 4, c8 = (v & 15) * 8;
      const bool ok = kc * LV_KC + c8 < d;      // the last stage of a width that is not a multiple of 128 is zero-filled
      cp_async_16(&Bs[st * LV_B + row * LV_LD + c8], b_base + row * ldw + (ok ? kc * LV_KC + c8 : 0), ok);
    }
  };
  auto load_a = [&](int kc, int st, int part) {
#pragma unroll
    for (int i = 0; i < 2 * LV_MT; ++i) {
      const int v = threadIdx.x + i * 128;
      const int row = v >> 4, c8 = (v & 15) * 8;
      const bool ok = m0 + row < R && kc * LV_KC + c8 < d;
      cp_async_16(&As[(part * NST + st) * LV_A + row * LV_LD + c8],
                  a_base + part * part_stride + (ok ? static_cast<long long>(row) * H * d + kc * LV_KC + c8 : 0), ok);
    }
  };
  // the weights do not depend on the previous kernel: request them before the dependency wait
  for (int s = 0; s < NST - 1; ++s)
    if (s < n_chunks) load_b(s, s);
  pdl_wait();
  bool second = false;
  if (SPLIT) {
    // blend weights of the two parts of every row: w_p = l_p 2^(m_p - max) / sum
    bool mine = false;
    if (threadIdx.x < LV_ROWS) {
      const int row = m0 + threadIdx.x;
      float w0 = 1.f, w1 = 0.f;
      if (row < R) {
        const float2 p0 = ml[static_cast<long long>(row) * 32 + h];
        const float2 p1 = ml[(static_cast<long long>(R) + row) * 32 + h];
        if (p1.y > 0.f) {
          const float mx = fmaxf(p0.x, p1.x);
          const float e0 = p0.y * ex2_approx(p0.x - mx), e1 = p1.y * ex2_approx(p1.x - mx);
          const float inv = 1.0f / (e0 + e1);
          w0 = e0 * inv;
          w1 = e1 * inv;
          mine = true;
        }
      }
      wgt[0][threadIdx.x] = w0;
      wgt[1][threadIdx.x] = w1;
    }
    second = __syncthreads_or(mine ? 1 : 0) != 0;
  }
  for (int s = 0; s < NST - 1; ++s) {
    if (s < n_chunks) {
      load_a(s, s, 0);
      if (SPLIT && second) load_a(s, s, 1);
    }
    cp_async_commit();     // group s = {A(s)} (+ all early B loads in group 0)
  }
  float c[NP][LV_MT][2][4];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int mt = 0; mt < LV_MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[p][mt][nt][e] = 0.f;
  for (int kc = 0; kc < n_chunks; ++kc) {
    cp_async_wait<NST - 2>();
    __syncthreads();
    const int nx = kc + NST - 1;
    if (nx < n_chunks) {
      load_b(nx, nx % NST);
      load_a(nx, nx % NST, 0);
      if (SPLIT && second) load_a(nx, nx % NST, 1);
    }
    cp_async_commit();
    const int st = kc % NST;
#pragma unroll
    for (int kp = 0; kp < LV_KC / 32; ++kp) {
      uint32_t bf[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        ldmatrix_x4(bf[nt], &Bs[st * LV_B + ((warp * 2 + nt) * 8 + (lane & 7)) * LV_LD + kp * 32 + (lane >> 3) * 8]);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (p == 1 && !second) continue;
#pragma unroll
        for (int mt = 0; mt < LV_MT; ++mt) {
          uint32_t a0[4], a1[4];
          const __nv_bfloat16* ap = &As[(p * NST + st) * LV_A + mt * 16 * LV_LD];
          ldmatrix_x4(a0, ap + (lane & 15) * LV_LD + kp * 32 + (lane >> 4) * 8);
          ldmatrix_x4(a1, ap + (lane & 15) * LV_LD + kp * 32 + 16 + (lane >> 4) * 8);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            mma_bf16_16816(c[p][mt][nt], a0, bf[nt][0], bf[nt][1]);
            mma_bf16_16816(c[p][mt][nt], a1, bf[nt][2], bf[nt][3]);
          }
        }
      }
    }
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < LV_MT; ++mt) {
    const int ra = mt * 16 + g, rb = ra + 8;                // rows of this thread inside the CTA's tile
    float wa0 = 1.f, wb0 = 1.f, wa1 = 0.f, wb1 = 0.f;       // blend weights of rows ra / rb, parts 0 / 1
    if (SPLIT && second) { wa0 = wgt[0][ra]; wb0 = wgt[0][rb]; wa1 = wgt[1][ra]; wb1 = wgt[1][rb]; }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = h * 64 + (warp * 2 + nt) * 8 + 2 * t;
      const float b0 = bv ? bv[col] : 0.f, b1 = bv ? bv[col + 1] : 0.f;
      float v0 = c[0][mt][nt][0], v1 = c[0][mt][nt][1], v2 = c[0][mt][nt][2], v3 = c[0][mt][nt][3];
      if (SPLIT && second) {
        v0 = wa0 * v0 + wa1 * c[NP - 1][mt][nt][0];
        v1 = wa0 * v1 + wa1 * c[NP - 1][mt][nt][1];
        v2 = wb0 * v2 + wb1 * c[NP - 1][mt][nt][2];
        v3 = wb0 * v3 + wb1 * c[NP - 1][mt][nt][3];
      }
      if (m0 + ra < R) *reinterpret_cast<uint32_t*>(o + (m0 + ra) * ldo + col) = pack_bf16(v0 + b0, v1 + b1);
      if (m0 + rb < R) *reinterpret_cast<uint32_t*>(o + (m0 + rb) * ldo + col) = pack_bf16(v2 + b0, v3 + b1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same projection with TMA-fed operand tiles.  latent_value_kernel moves its 240 - 330 KB per CTA with 16-byte
// cp.async: ~15 B/clk per SM whatever the ring depth (3 and 6 stages: 7.3 / 11.8 us both) - the LSU path, not latency,
// bounds it.  Here a fifth warp feeds a ring of 128-column stages with TMA boxes (128-byte swizzle, read back by
// ldmatrix through the same XOR), the weight boxes of the first stages before the dependency wait; the four MMA warps
// only wait on mbarriers.  grid (H, ceil(R / 32)), 160 threads; needs d % 128 == 0 (always true on the latent path: it
// takes an even number of 64-wide heads); WF_LATENT_VALUE_TMA=0 selects the cp.async kernel for A/B runs.
static constexpr int LVT_STAGES = 6;
static constexpr int LVT_ROWS = 32;
static constexpr int LVT_A_ATOM = LVT_ROWS * 128;      // 32 rows x 64 columns, bytes
static constexpr int LVT_B_ATOM = 64 * 128;            // 64 output columns x 64 columns of K, bytes

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(160)
latent_value_tma_kernel(const __grid_constant__ CUtensorMap ma0, const __grid_constant__ CUtensorMap ma1,
                        const __grid_constant__ CUtensorMap mb, const float2* __restrict__ ml,
                        const float* __restrict__ bv, __nv_bfloat16* __restrict__ o, long long ldo, int R, int H, int d,
                        int early) {
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int A_STAGE = 2 * LVT_A_ATOM;                 // per part
  constexpr int STAGE = NP * A_STAGE + 2 * LVT_B_ATOM;    // 24 KB, 32 KB in the split form
  extern __shared__ __align__(1024) uint8_t lvt_smem[];
  __shared__ __align__(8) uint64_t full[LVT_STAGES], empty[LVT_STAGES];
  __shared__ float wgt[2][LVT_ROWS];
  __shared__ int second_flag;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lvt_smem) + 1023) & ~static_cast<uintptr_t>(1023));
  const int h = blockIdx.x, m0 = blockIdx.y * LVT_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = d / 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LVT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    mbar_fence_init();
  }
  __syncthreads();
  if (early) pdl_trigger();   // the out-projection GEMM behind this kernel may place its CTAs and request its weights

  if (warp == 4) {   // ---- producer
    if (lane == 0) {
      tma_prefetch_desc(&ma0);
      tma_prefetch_desc(&mb);
      const int pre = n_chunks < LVT_STAGES ? n_chunks : LVT_STAGES;
      for (int s = 0; s < pre; ++s) {   // the weights do not depend on the previous kernel
        uint8_t* st = ring + s * STAGE + NP * A_STAGE;
        mbar_expect_tx(&full[s], 2 * LVT_B_ATOM);
        tma_load_2d(st, &mb, &full[s], s * 128, h * 64);
        tma_load_2d(st + LVT_B_ATOM, &mb, &full[s], s * 128 + 64, h * 64);
      }
    }
    pdl_wait();
    bool second = false;
    if (SPLIT) {
      const int row = m0 + lane;
      const bool mine = row < R && ml[(static_cast<long long>(R) + row) * 32 + h].y > 0.f;
      second = __any_sync(0xffffffffu, mine);
    }
    if (lane == 0) {
      for (int kc = 0; kc < n_chunks; ++kc) {
        const int s = kc % LVT_STAGES;
        uint8_t* st = ring + s * STAGE;
        if (kc >= LVT_STAGES) {
          mbar_wait(&empty[s], ((kc / LVT_STAGES) - 1) & 1);
          mbar_expect_tx(&full[s], 2 * LVT_B_ATOM);
          tma_load_2d(st + NP * A_STAGE, &mb, &full[s], kc * 128, h * 64);
          tma_load_2d(st + NP * A_STAGE + LVT_B_ATOM, &mb, &full[s], kc * 128 + 64, h * 64);
        }
        mbar_arrive_expect_tx(&full[s], (SPLIT && second ? 2 : 1) * A_STAGE);
        tma_load_2d(st, &ma0, &full[s], h * d + kc * 128, m0);
        tma_load_2d(st + LVT_A_ATOM, &ma0, &full[s], h * d + kc * 128 + 64, m0);
        if (SPLIT && second) {
          tma_load_2d(st + A_STAGE, &ma1, &full[s], h * d + kc * 128, m0);
          tma_load_2d(st + A_STAGE + LVT_A_ATOM, &ma1, &full[s], h * d + kc * 128 + 64, m0);
        }
      }
    }
    return;
  }

  // ---- consumers: warp w owns output columns 16 w .. 16 w + 15 of the head
  pdl_wait();
  bool second = false;
  if (SPLIT) {
    if (warp == 0) {   // blend weights of the two parts of every row: w_p = l_p 2^(m_p - max) / sum
      const int row = m0 + lane;
      float w0 = 1.f, w1 = 0.f;
      bool mine = false;
      if (row < R) {
        const float2 p0 = ml[static_cast<long long>(row) * 32 + h];
        const float2 p1 = ml[(static_cast<long long>(R) + row) * 32 + h];
        if (p1.y > 0.f) {
          const float mx = fmaxf(p0.x, p1.x);
          const float e0 = p0.y * ex2_approx(p0.x - mx), e1 = p1.y * ex2_approx(p1.x - mx);
          const float inv = 1.0f / (e0 + e1);
          w0 = e0 * inv;
          w1 = e1 * inv;
          mine = true;
        }
      }
      wgt[0][lane] = w0;
      wgt[1][lane] = w1;
      const bool any = __any_sync(0xffffffffu, mine);
      if (lane == 0) second_flag = any ? 1 : 0;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    second = second_flag != 0;
  }
  float c[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[mt][nt][e] = 0.f;
  const uint32_t ring_u = smem_u32(ring);
  // swizzled byte offsets inside an atom: row r, 16-byte chunk ch -> r * 128 + ((ch ^ (r & 7)) << 4)
  const int a_row = lane & 15, a_ch = lane >> 4;
  const int b_ch = lane >> 3;
  int b_row[2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) b_row[nt] = (warp * 2 + nt) * 8 + (lane & 7);
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int s = kc % LVT_STAGES;
    mbar_wait(&full[s], (kc / LVT_STAGES) & 1);
    const uint32_t st = ring_u + s * STAGE;
    if (SPLIT && second) {
      // The projection is linear in the context: blend the two parts in place (w0 c0 + w1 c1, one more bf16 rounding of
      // the context) and project once - a second set of mma.sync costs 3 us per launch, the blend 8 FMAs per 16 bytes.
      uint8_t* a0p = ring + s * STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = (threadIdx.x + i * 128);          // 16-byte chunk of the 2 x 4 KB of part 0 (same spot in part 1)
        const int row = (ch & 255) >> 3;
        const float w0 = wgt[0][row], w1 = wgt[1][row];
        uint4 x = *reinterpret_cast<const uint4*>(a0p + ch * 16);
        const uint4 y = *reinterpret_cast<const uint4*>(a0p + A_STAGE + ch * 16);
        x.x = pack_bf16(w0 * bf16lo(x.x) + w1 * bf16lo(y.x), w0 * bf16hi(x.x) + w1 * bf16hi(y.x));
        x.y = pack_bf16(w0 * bf16lo(x.y) + w1 * bf16lo(y.y), w0 * bf16hi(x.y) + w1 * bf16hi(y.y));
        x.z = pack_bf16(w0 * bf16lo(x.z) + w1 * bf16lo(y.z), w0 * bf16hi(x.z) + w1 * bf16hi(y.z));
        x.w = pack_bf16(w0 * bf16lo(x.w) + w1 * bf16lo(y.w), w0 * bf16hi(x.w) + w1 * bf16hi(y.w));
        *reinterpret_cast<uint4*>(a0p + ch * 16) = x;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {   // 32 columns of K each; atom = kp / 2
      const int ch0 = (kp & 1) * 4;
      const uint32_t b_atom = st + NP * A_STAGE + (kp >> 1) * LVT_B_ATOM;
      uint32_t bf[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t addr = b_atom + b_row[nt] * 128 + (((ch0 + b_ch) ^ (b_row[nt] & 7)) << 4);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bf[nt][0]), "=r"(bf[nt][1]), "=r"(bf[nt][2]), "=r"(bf[nt][3]) : "r"(addr));
      }
      {
        const uint32_t a_atom = st + (kp >> 1) * LVT_A_ATOM;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r = mt * 16 + a_row;
          uint32_t a0[4], a1[4];
          const uint32_t ad0 = a_atom + r * 128 + (((ch0 + a_ch) ^ (r & 7)) << 4);
          const uint32_t ad1 = a_atom + r * 128 + (((ch0 + 2 + a_ch) ^ (r & 7)) << 4);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]) : "r"(ad0));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]) : "r"(ad1));
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            mma_bf16_16816(c[mt][nt], a0, bf[nt][0], bf[nt][1]);
            mma_bf16_16816(c[mt][nt], a1, bf[nt][2], bf[nt][3]);
          }
        }
      }
    }
    if (SPLIT && second) fence_proxy_async_smem();   // the blend wrote the stage the next TMA box overwrites
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int ra = mt * 16 + g, rb = ra + 8;                // rows of this thread inside the CTA's tile
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = h * 64 + (warp * 2 + nt) * 8 + 2 * t;
      const float b0 = bv ? bv[col] : 0.f, b1 = bv ? bv[col + 1] : 0.f;
      const float v0 = c[mt][nt][0], v1 = c[mt][nt][1], v2 = c[mt][nt][2], v3 = c[mt][nt][3];
      if (m0 + ra < R) *reinterpret_cast<uint32_t*>(o + (m0 + ra) * ldo + col) = pack_bf16(v0 + b0, v1 + b1);
      if (m0 + rb < R) *reinterpret_cast<uint32_t*>(o + (m0 + rb) * ldo + col) = pack_bf16(v2 + b0, v3 + b1);
    }
  }
}

int latent_value(const void* ctx, long long part_stride, const float* ml, const void* wv, long long ldw, const float* bv,
                 void* o, long long ldo, int R, int H, cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(ldw % 8 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(wv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 3) == 0 &&
                 part_stride % 8 == 0,
             "latent value: operands must be 16-byte aligned");
  WF_REQUIRE(H <= 32, "latent value: at most 32 heads");
  static const bool no_tma = getenv("WF_LATENT_VALUE_TMA") != nullptr && atoi(getenv("WF_LATENT_VALUE_TMA")) == 0;
  if (d % 128 == 0 && ldw % 8 == 0 && !no_tma) {
    const bool split = ml != nullptr;
    CUtensorMap ma0, ma1, mb;
    int rc = make_map_bf16(&ma0, ctx, R, static_cast<long long>(H) * d, static_cast<long long>(H) * d, LVT_ROWS);
    if (rc != WF_OK) return rc;
    ma1 = ma0;
    if (split) {
      rc = make_map_bf16(&ma1, reinterpret_cast<const __nv_bfloat16*>(ctx) + part_stride, R, static_cast<long long>(H) * d,
                         static_cast<long long>(H) * d, LVT_ROWS);
      if (rc != WF_OK) return rc;
    }
    rc = make_map_bf16(&mb, wv, d, d, ldw, 64);
    if (rc != WF_OK) return rc;
    const size_t stage = (split ? 2 : 1) * 2 * LVT_A_ATOM + 2 * LVT_B_ATOM;
    const size_t smem = LVT_STAGES * stage + 1024;
    static PerDeviceOnce tconf;
    if (tconf.first_use()) {
      const int big = LVT_STAGES * (4 * LVT_A_ATOM + 2 * LVT_B_ATOM) + 1024;
      WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    }
    dim3 tgrid(H, (R + LVT_ROWS - 1) / LVT_ROWS);
    static const int early = getenv("WF_LV_EARLY") ? atoi(getenv("WF_LV_EARLY")) : 1;   // measured: decode loop -1.3 .. -2.0 ms per step
    if (split)
      WF_CHECK_CUDA(launch_pdl(0, latent_value_tma_kernel<true>, tgrid, dim3(160), smem, stream, ma0, ma1, mb,
                               reinterpret_cast<const float2*>(ml), bv, reinterpret_cast<__nv_bfloat16*>(o), ldo, R, H, d, early));
    else
      WF_CHECK_CUDA(launch_pdl(0, latent_value_tma_kernel<false>, tgrid, dim3(160), smem, stream, ma0, ma1, mb,
                               reinterpret_cast<const float2*>(ml), bv, reinterpret_cast<__nv_bfloat16*>(o), ldo, R, H, d, early));
    count_launch();
    return WF_OK;
  }
  const bool wide = H * ((R + 15) / 16) > num_sms();      // 32 rows per CTA
  const int rows = wide ? 32 : 16;
  dim3 grid(H, (R + rows - 1) / rows);
  const size_t smem1 = static_cast<size_t>(LV_STAGES) * ((rows / 16) * LV_A16 + LV_B) * 2;
  const size_t smem2 = smem1 + static_cast<size_t>(LV_STAGES) * (rows / 16) * LV_A16 * 2;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    const int big = static_cast<int>(static_cast<size_t>(LV_STAGES) * (4 * LV_A16 + LV_B) * 2);
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_value_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  }
  const __nv_bfloat16* c16 = reinterpret_cast<const __nv_bfloat16*>(ctx);
  const __nv_bfloat16* w16 = reinterpret_cast<const __nv_bfloat16*>(wv);
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(o);
  const float2* ml2 = reinterpret_cast<const float2*>(ml);
  if (ml != nullptr) {
    if (wide) WF_CHECK_CUDA(launch_pdl(0, latent_value_kernel<true, 2>, grid, dim3(128), smem2, stream, c16, part_stride, ml2,
                                       w16, ldw, bv, o16, ldo, R, H, d));
    else WF_CHECK_CUDA(launch_pdl(0, latent_value_kernel<true, 1>, grid, dim3(128), smem2, stream, c16, part_stride, ml2, w16,
                                  ldw, bv, o16, ldo, R, H, d));
  } else {
    if (wide) WF_CHECK_CUDA(launch_pdl(0, latent_value_kernel<false, 2>, grid, dim3(128), smem1, stream, c16, 0LL, ml2, w16,
                                       ldw, bv, o16, ldo, R, H, d));
    else WF_CHECK_CUDA(launch_pdl(0, latent_value_kernel<false, 1>, grid, dim3(128), smem1, stream, c16, 0LL, ml2, w16, ldw,
                                  bv, o16, ldo, R, H, d));
  }
  count_launch();
  return WF_OK;
}

}  // namespace wf
