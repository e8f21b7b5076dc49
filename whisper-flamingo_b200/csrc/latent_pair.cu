// One-token cross-attention over the source rows in ONE pass: a thread-block cluster of two CTAs owns a clip, each CTA
// holds half of the latent width d, and every 64-key tile stays in shared memory between the score product and the
// context product.  Same contract as latent_attn_kernel (latent.cu): q' = Wk_h^T q_h, c_h = softmax(src q'_h / 8)^T src.
// Replaces the per-step recompute of reference whisper/decoding.py:155-164 for the cross-attention (model.py:93-108
// with xa) and the gated x-attention (model.py:110-134 with xt).
//
// Why: the two-pass kernel of latent.cu reads every tile twice (HBM, then L2).  Its issuing threads wait ~1000 clk per
// 19 KB chunk: 128 CTAs x ~51 B/clk = 6500 B/clk, which is what the L2 slices of the chip deliver (~12 TB/s, only 2x
// HBM) - it is L2-bound at 114-118 us = 61-65 % of the HBM roofline, and more issuing threads, more SMs or deeper
// rings do not move it (profiles/r02_latent_experiments.txt).  One pass needs the tile on chip: 64 keys x d = 160 KB at
// d = 1280 leaves room for one tile per SM, and a single-tile ring serialises load latency, both products and the
// softmax hand-off of every tile (measured 218 us, tools/probe/latent_onepass.cu).  Split over two SMs a tile is
// 80 KB per CTA, the ring holds two, and the products of consecutive tiles overlap.
//   CTA r of the pair: columns [r d/2, (r + 1) d/2) of the source rows and of q' (resident, 30 KB).
//   pass A  S_r^T[64 keys x 32 heads] += stage (A, K-major, 64 keys x 128 columns) x q'^T (B), K = d / 2
//   exchange each softmax thread (= key) sends its partial scores to the peer through distributed shared memory
//           (5 KB per tile and direction) and adds the peer's: both CTAs hold the same S^T, bit for bit, and keep the
//           same running reference maxima and sums without further communication
//   softmax as in latent.cu (reference maximum with a 2^8 window), P^T (bf16) -> smem of the own CTA
//   pass B  C^T[128 columns x 32 heads] += stage^T (A, MN-major: the same swizzled bytes) x P^T (B); its commit frees
//           the stage for the next tile
// HBM traffic: every source row once; L2 -> SM traffic: the same bytes.  MMA count per CTA and 64 keys: 40 + 20 narrow
// tcgen05.mma.
// Scheduling: one cluster per clip is 256 CTAs for 128 clips = 1.73 waves on 148 SMs (measured 119.8 us, a single
// wave of 64 clips 49.8 us).  With a partial-state buffer (`ml`) the kernel is persistent instead: num_sms / 2
// clusters take equal contiguous ranges of the B x n_tiles tile sequence; a range that ends inside a clip leaves that
// clip in two segments, each written as its own normalised context (part 0 = the segment that starts the clip, part 1
// = the rest) with its (reference maximum, row sum) per head.  The value projection that follows is linear in the
// context, so latent_value_kernel projects both parts and blends them with the softmax weights of the two segments -
// no communication between clusters and no co-residency requirement.  (Blending inside this kernel - the cluster
// that owns the head of a cut clip waits for a flag from its neighbour - was built and measured: 111 us instead of
// 102.5, the blend being the last thing a cluster does, and two such kernels on concurrent streams could each hold
// SMs while waiting for clusters that cannot be scheduled.)  The ring keeps streaming across segment borders; only q'
// is reloaded (one L2 round trip) and the context accumulators alternate between two TMEM regions.
// What bounds it (profiles/r02_latent_experiments.txt): a ring stage is occupied for (load latency under load ~2.4 us)
// + (residency: scores of the whole tile, exchange + softmax, context MMAs ~2.3 us), so 160 KB of ring per SM turn
// over at ~17 B/clk/SM = 4.8 TB/s.  With the MMAs and the softmax removed the same ring streams the same bytes in
// 75.0 us = 6.56 TB/s: the memory path is not the limit, the loop is.  Measured dead ends: pass A yielding the tensor
// pipe to pass B (118 us vs 102.5), a tile-contiguous source layout (no change: DRAM page locality is not the limit),
// one issuer per pass (108 us), L2 prefetch 1 .. 4 tiles ahead of the ring (105.9 .. 111.9 us; it slows even the bare
// stream, 84 us).
#include "common.cuh"
#include "kernels.h"

namespace wf {

static constexpr int LP_KT = 64;                    // keys per tile
static constexpr int LP_BOX = LP_KT * 128;          // 8 KB: [64 keys x 64 columns] bf16, 128B-swizzled = one TMA box
static constexpr int LP_STAGE = 2 * LP_BOX;         // 16 KB: 128 columns = the A operand of one context accumulator
static constexpr int LP_NH = 32;                    // head columns of both MMAs (H <= 32)
static constexpr int LP_PT = LP_NH * 128;           // 4 KB: P^T rows (heads) x 64 keys
static constexpr int LP_XLD = 36;                   // floats per key row of the exchange buffer (144 B: 16-byte vectors of
                                                    // neighbouring keys fall into different banks)
static constexpr int LP_XCH = LP_KT * LP_XLD * 4;   // 9 KB: partial scores of the peer, [key][head] fp32
static constexpr int LP_MISC = 1536;
static constexpr int LP_MAX_S = 14;
static constexpr int LP_SMEM_LIMIT = 227 * 1024;
#ifndef LP_ISSUERS
#define LP_ISSUERS 2       // MMA-issuing threads per pass (own accumulators): a narrow tcgen05.mma occupies the tensor front
#endif                     // end for 40 clk but costs its issuing thread 53-66 clk, and a stage sits in the ring until
                           // the scores of its whole tile, the softmax and its context MMAs are through - the ring
                           // turns over in (load latency + residency), so shorter issue phases are bandwidth
static constexpr int LP_NI = LP_ISSUERS;
static_assert(LP_NI == 1 || LP_NI == 2, "one or two issuing threads per pass");
static constexpr int LP_TMEM_S = 2 * LP_NI * LP_NH;  // S^T double-buffered (x issuers), then 2 x d / 256 context accumulators
static constexpr int LP_TMEM_COLS = 512;
static constexpr int LP_THREADS = 256 + (LP_NI - 1) * 64;   // warps 8, 9: second issuer of pass A / pass B

__device__ __forceinline__ void lp_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void lp_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// MN-major operand, 128B swizzle: rows = K index (keys; 128 B = 64 MN elements each), 8-row groups 1024 B apart (SBO),
// the next 64 MN elements one TMA box further (LBO)
__device__ __forceinline__ uint64_t lp_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(LP_BOX >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t lp_idesc_a_mn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 15); }

__device__ __forceinline__ uint32_t lp_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void lp_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t lp_mapa(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
// 16 bytes into the peer's shared memory; the store itself credits the peer's mbarrier (complete_tx), so the receiver
// needs no cluster-scope fence: its mbarrier wait makes the data visible (release.cluster arrivals + acquire.cluster
// waits were measured first: ~2700 clk per tile and thread)
__device__ __forceinline__ void lp_st_async_v4(uint32_t addr, uint32_t bar_cluster_addr, uint32_t a, uint32_t b, uint32_t c,
                                               uint32_t d) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];"
               ::"r"(addr), "r"(bar_cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// Tile sequence of a cluster: global tiles [g_lo, g_hi) cut at clip borders into segments (clip, [t_lo, t_hi)).
struct LpSegs {
  long long g, g_hi;
  int n_tiles;
  int clip, t_lo, t_hi;
  __device__ LpSegs(long long lo, long long hi, int nt) : g(lo), g_hi(hi), n_tiles(nt), clip(0), t_lo(0), t_hi(0) {}
  __device__ bool next() {
    if (g >= g_hi) return false;
    clip = static_cast<int>(g / n_tiles);
    t_lo = static_cast<int>(g - static_cast<long long>(clip) * n_tiles);
    const long long left = g_hi - g;
    t_hi = left < n_tiles - t_lo ? t_lo + static_cast<int>(left) : n_tiles;
    g += t_hi - t_lo;
    return true;
  }
};

// grid = 2 x clusters (cluster of 2); NS = d / 256 stages per tile and CTA; NSLOT ring stages (>= 2 NS).
// ml == nullptr: one cluster per clip (gridDim.x == 2 B), context in ctx.  Otherwise the clusters share the tile
// sequence evenly; ctx holds two parts `part_stride` elements apart, ml[(part * B + clip) * 32 + head] = (reference
// maximum in log2 units, row sum) of that part (row sum 0: the part is absent).
// HT = number of heads at compile time (0: the run-time H).  The softmax warps run one per scheduler, so their
// instruction count is their time: with a run-time bound the 32-way unrolled head loops carry a predicate per head
// (~30 instructions per head, 5500 of the 7700 clk of a tile); with HT they are straight-line code.
template <int HT>
__global__ void __launch_bounds__(LP_THREADS, 1)
latent_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q,
                   __nv_bfloat16* __restrict__ ctx, float2* __restrict__ ml, long long part_stride, int B, int T, int H,
                   int HP, int NS, int NSLOT, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_atom = HP * 128;
  uint8_t* ring = smem;                              // NSLOT stages
  uint8_t* qs = ring + NSLOT * LP_STAGE;             // q' (own column half): 2 NS K-major atoms of HP rows (heads)
  uint8_t* pt = qs + 2 * NS * q_atom;                // P^T operand, double-buffered
  float* xch = reinterpret_cast<float*>(pt + 2 * LP_PT);   // [2][key][head] partial scores written by the peer
  uint8_t* misc = reinterpret_cast<uint8_t*>(xch) + 2 * LP_XCH;
  float* m_buf = reinterpret_cast<float*>(misc);     // [32] reference maximum of each head
  float* al_buf = m_buf + 32;                        // [32] rescale factor when the reference moved
  float* linv_buf = al_buf + 32;                     // [32]
  float* red = linv_buf + 32;                        // [4][32]
  int* flag_buf = reinterpret_cast<int*>(red + 128); // [2][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 944);
  uint64_t* full_s = bars;                           // [NSLOT] stage landed
  uint64_t* empty_s = full_s + LP_MAX_S;             // [NSLOT] context MMAs of the stage completed
  uint64_t* q_full = empty_s + LP_MAX_S;             // q' of a segment landed
  uint64_t* q_free = q_full + 1;                     // the score MMAs of a segment have read q' (one commit per issuer)
  uint64_t* s_full = q_free + 1;                     // [2] partial scores of a tile in TMEM
  uint64_t* s_free = s_full + 2;                     // [2] ... copied to registers (128 arrivals)
  uint64_t* x_full = s_free + 2;                     // [2] the peer's partial scores landed in xch (transaction bytes)
  uint64_t* p_ready = x_full + 2;                    // [2] P^T staged, C^T rescaled (128 arrivals)
  uint64_t* c_done = p_ready + 2;                    // [2] context MMAs of a tile completed
  uint64_t* c_free = c_done + 2;                     // [2] context accumulators of a segment copied out (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hc = HT ? HT : H;                        // head count the unrolled loops are bounded by
  const uint32_t rank = lp_ctarank();
  const int n_tiles = (T + LP_KT - 1) / LP_KT;
  const int d = NS * 256;
  const int col0 = static_cast<int>(rank) * (d >> 1);   // first latent column of this CTA
  const long long total = static_cast<long long>(B) * n_tiles;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const long long g_lo = total * cid / ncl, g_hi = total * (cid + 1) / ncl;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_q);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full_s[i], 1); mbar_init(&empty_s[i], 1); }
    mbar_init(q_full, 1);
    mbar_init(q_free, LP_NI);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], LP_NI); mbar_init(&s_free[i], 128); mbar_init(&x_full[i], 1);
      mbar_init(&p_ready[i], 128); mbar_init(&c_done[i], LP_NI); mbar_init(&c_free[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<LP_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  lp_cluster_sync();      // the peer's barriers exist before anything arrives on them

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer: source stages (static data: the
    // first NSLOT stages are requested before the previous kernel has finished).  The ring does not see segment borders.
    const uint64_t once = l2_policy_evict_first();
    int slot = 0;
    uint32_t phase = 0;
#ifdef LP_TIMING
    long long tp_w = 0, tp_n = 0;
#endif
    LpSegs sg(g_lo, g_hi, n_tiles);
    while (sg.next()) {
      for (int j = sg.t_lo; j < sg.t_hi; ++j) {
        const int row = sg.clip * T + j * LP_KT;
        for (int a = 0; a < NS; ++a) {
#ifdef LP_TIMING
          const long long tp0 = clock64();
#endif
          mbar_wait(&empty_s[slot], phase ^ 1);
#ifdef LP_TIMING
          tp_w += clock64() - tp0; ++tp_n;
#endif
          mbar_arrive_expect_tx(&full_s[slot], LP_STAGE);
          uint8_t* dst = ring + slot * LP_STAGE;
          lp_tma_load_2d(dst, &map_x, &full_s[slot], col0 + a * 128, row, once);
          lp_tma_load_2d(dst + LP_BOX, &map_x, &full_s[slot], col0 + a * 128 + 64, row, once);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        }
      }
    }
#ifdef LP_TIMING
    if (blockIdx.x == 0) printf("producer: wait for a free slot %lld clk per stage (%lld stages)\n", tp_w / tp_n, tp_n);
#endif
  } else if ((warp == 1 || (LP_NI == 2 && warp == 8)) && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer(s), pass A (+ the q' of each segment)
    // S_r^T[64 keys x 32 heads] = stage (A, K-major) x q'^T (B, K-major; rows >= HP of an atom alias what follows it
    // and only produce head columns nobody reads), K = d / 2 in 16-column steps.  Issuer `me` takes the stages whose
    // ring counter is = me (mod LP_NI) - always the same ring slots, the ring is even - into its own accumulator.
    constexpr uint32_t idesc_s = umma_idesc_bf16(LP_KT, LP_NH);
    const uint32_t ra = smem_u32(ring), qa = smem_u32(qs);
    const int me = warp == 1 ? 0 : 1;
    int it = 0, seg = 0, slot = 0;
    uint32_t phase = 0, gs = 0;      // ring slot / phase / parity of the stage counter
#ifdef LP_TIMING
    long long ta_q = 0, ta_sf = 0, ta_w = 0, ta_m = 0;
#endif
    pdl_wait();       // q' comes from the previous kernel
    LpSegs sg(g_lo, g_hi, n_tiles);
    while (sg.next()) {
      if (me == 0) {
        // every score MMA of the previous segment has read the old q' before it is replaced
        if (seg > 0) mbar_wait(q_free, (seg - 1) & 1);
        mbar_arrive_expect_tx(q_full, 2 * NS * q_atom);
        for (int i = 0; i < 2 * NS; ++i) tma_load_2d(qs + i * q_atom, &map_q, q_full, col0 + i * 64, sg.clip * H);
      }
#ifdef LP_TIMING
      long long tla = clock64();
#define LP_TA(acc_) do { const long long n_ = clock64(); acc_ += n_ - tla; tla = n_; } while (0)
#else
#define LP_TA(acc_)
#endif
      mbar_wait(q_full, seg & 1);
      LP_TA(ta_q);
      for (int j = sg.t_lo; j < sg.t_hi; ++j, ++it) {
        if (it >= 2) mbar_wait(&s_free[it & 1], ((it >> 1) - 1) & 1);
        LP_TA(ta_sf);
        tc_fence_after();
        const uint32_t acc = tmem_base + ((it & 1) * LP_NI + me) * LP_NH;
        bool first = true;
        for (int a = 0; a < NS; ++a, ++gs, slot = slot + 1 == NSLOT ? 0 : slot + 1, phase ^= (slot == 0)) {
          if (LP_NI == 2 && (gs & 1) != me) continue;
          mbar_wait(&full_s[slot], phase);
          LP_TA(ta_w);
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t a_desc = umma_desc_kmajor_sw128(ra + slot * LP_STAGE + half * LP_BOX);
            const uint64_t b_desc = umma_desc_kmajor_sw128(qa + (2 * a + half) * q_atom);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(acc, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (!first || half || k) ? 1u : 0u);
          }
          first = false;
          LP_TA(ta_m);
        }
        umma_commit(&s_full[it & 1]);
      }
      umma_commit(q_free);
      ++seg;
    }
#ifdef LP_TIMING
    if (blockIdx.x == 0)
      printf("issuer A%d per tile: wait q' %lld | wait s_free %lld | wait stages %lld | issue %lld ; tiles %d\n", me,
             ta_q / it, ta_sf / it, ta_w / it, ta_m / it, it);
#endif
  } else if ((warp == 3 || (LP_NI == 2 && warp == 9)) && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer(s), pass B
    // C^T[128 columns x 32 heads] (+)= stage^T (A, MN-major: rows = keys, two 64-column atoms 8 KB apart) x P^T (B,
    // K-major, one atom of 64 keys), K = 64 keys in 16-key steps.  The stage was waited for by pass A of the same tile.
    // The accumulators of consecutive segments alternate between two TMEM regions; issuer `me` takes the stages whose
    // ring counter is = me (mod LP_NI), every stage has its own accumulator.
    constexpr uint32_t idesc_c = lp_idesc_a_mn(128, LP_NH);
    const uint32_t ra = smem_u32(ring), pa = smem_u32(pt);
    const int me = warp == 3 ? 0 : 1;
    int it = 0, seg = 0, slot = 0;
    uint32_t gs = 0;
#ifdef LP_TIMING
    long long tb_p = 0, tb_m = 0;
#endif
    LpSegs sg(g_lo, g_hi, n_tiles);
    while (sg.next()) {
      const uint32_t acc = tmem_base + LP_TMEM_S + (seg & 1) * NS * LP_NH;
      if (seg >= 2) {
        mbar_wait(&c_free[seg & 1], ((seg >> 1) - 1) & 1);
        tc_fence_after();
      }
      for (int j = sg.t_lo; j < sg.t_hi; ++j, ++it) {
#ifdef LP_TIMING
        const long long tb0 = clock64();
#endif
        mbar_wait(&p_ready[it & 1], (it >> 1) & 1);
#ifdef LP_TIMING
        const long long tb1 = clock64();
        tb_p += tb1 - tb0;
#endif
        tc_fence_after();
        const uint64_t p_desc = umma_desc_kmajor_sw128(pa + (it & 1) * LP_PT);
        for (int a = 0; a < NS; ++a, ++gs, slot = slot + 1 == NSLOT ? 0 : slot + 1) {
          if (LP_NI == 2 && (gs & 1) != me) continue;
          const uint32_t st = ra + slot * LP_STAGE;
#pragma unroll
          for (int kk = 0; kk < LP_KT / 16; ++kk)
            umma_f16(acc + a * LP_NH, lp_desc_mn(st + kk * 2048), p_desc + 2 * kk, idesc_c,
                     (j > sg.t_lo || kk > 0) ? 1u : 0u);
          umma_commit(&empty_s[slot]);
        }
        umma_commit(&c_done[it & 1]);
#ifdef LP_TIMING
        tb_m += clock64() - tb1;
#endif
      }
      ++seg;
    }
#ifdef LP_TIMING
    if (blockIdx.x == 0) printf("issuer B%d per tile: wait p_ready %lld | issue %lld\n", me, tb_p / it, tb_m / it);
#endif
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ softmax (thread = key = TMEM lane of S^T)
    const int wq = warp - 4;
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float l_part[LP_NH], msl[LP_NH];                   // row-sum partials; reference maxima in log2 units (m * sl2)
    pdl_wait();     // the context rows are read by an earlier kernel of the stream
    // a 64-row accumulator keeps rows 16 q .. 16 q + 15 in lanes 32 q .. 32 q + 15 (profiles/r01_probe_tmem_m64_layout.txt)
    const int key = wq * 16 + (lane & 15);
    const bool lane_on = lane < 16;
    // P^T element (head h, my key): row = head (128 B), 16-byte units swizzled by the row
    const uint32_t p_off = (key & 7) * 2;
    const uint32_t p_unit = key >> 3;
    const uint32_t peer = rank ^ 1u;
    const uint32_t xch_peer = lp_mapa(smem_u32(xch), peer);
    const uint32_t xfull_peer = lp_mapa(smem_u32(x_full), peer);
    const int hv = (Hc + 3) >> 2;                      // 16-byte vectors of head scores per key
    int it = 0, seg = 0;
#ifdef LP_TIMING
    long long t_sf = 0, t_st = 0, t_xw = 0, t_mid = 0, t_pw = 0, t_p = 0, t_epi = 0;
    const long long tt0 = clock64();
#endif
    LpSegs sg(g_lo, g_hi, n_tiles);
    while (sg.next()) {
      const uint32_t acc = lane_base + LP_TMEM_S + (seg & 1) * NS * LP_NH;
#pragma unroll
      for (int h = 0; h < LP_NH; ++h) { l_part[h] = 0.f; msl[h] = -INFINITY; }
      if (tid < LP_NH) m_buf[tid] = -INFINITY;
      lp_bar(1);
      for (int j = sg.t_lo; j < sg.t_hi; ++j, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
#ifdef LP_TIMING
        long long tl = clock64();
#define LP_T(acc_) do { const long long n_ = clock64(); acc_ += n_ - tl; tl = n_; } while (0)
#else
#define LP_T(acc_)
#endif
        mbar_wait(&s_full[buf], ph);
        LP_T(t_sf);
        tc_fence_after();
        uint32_t sv[32];
        tmem_ld_32x32(lane_base + buf * LP_NI * LP_NH, sv);
        if (LP_NI == 2) {    // the two issuers of pass A dealt the stages to two accumulators
          uint32_t sv2[32];
          tmem_ld_32x32(lane_base + (buf * LP_NI + 1) * LP_NH, sv2);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < LP_NH; ++h)
            if (h < 4 * hv) sv[h] = __float_as_uint(__uint_as_float(sv[h]) + __uint_as_float(sv2[h]));
        } else {
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&s_free[buf]);
        // partial scores -> the peer's exchange buffer [buf][key][head]; its previous content (tile it - 2) was consumed
        // before the peer's threads sent me tile it - 1, which I waited for before I got here
        if (tid == 0) mbar_arrive_expect_tx(&x_full[buf], static_cast<uint32_t>(LP_KT * hv * 16));
        if (lane_on) {
          const uint32_t dst = xch_peer + static_cast<uint32_t>(((buf * LP_KT + key) * LP_XLD) * 4);
#pragma unroll
          for (int v = 0; v < LP_NH / 4; ++v)
            if (v < hv)
              lp_st_async_v4(dst + v * 16, xfull_peer + buf * 8, sv[4 * v], sv[4 * v + 1], sv[4 * v + 2], sv[4 * v + 3]);
        }
        LP_T(t_st);
        mbar_wait(&x_full[buf], ph);
        LP_T(t_xw);
        // scores relative to the reference maximum of their head, in log2 units: t = (own + peer) * sl2 - m * sl2
        float t[LP_NH];
        {
          const float4* src = reinterpret_cast<const float4*>(xch + (buf * LP_KT + (lane_on ? key : 0)) * LP_XLD);
#pragma unroll
          for (int v = 0; v < LP_NH / 4; ++v)
            if (v < hv) {
              const float4 o = src[v];
              sv[4 * v] = __float_as_uint(__uint_as_float(sv[4 * v]) + o.x);
              sv[4 * v + 1] = __float_as_uint(__uint_as_float(sv[4 * v + 1]) + o.y);
              sv[4 * v + 2] = __float_as_uint(__uint_as_float(sv[4 * v + 2]) + o.z);
              sv[4 * v + 3] = __float_as_uint(__uint_as_float(sv[4 * v + 3]) + o.w);
            }
        }
        const bool valid = lane_on && j * LP_KT + key < T;
        // does any score leave the 2^8 window above its head's reference?  (always on the first tile: m = -inf)
        bool exceed = false;
#pragma unroll
        for (int h = 0; h < LP_NH; ++h)
          if (h < Hc) {
            t[h] = fmaf(__uint_as_float(sv[h]), sl2, -msl[h]);
            exceed |= t[h] > 8.f;
          }
        const bool w_any = __any_sync(0xffffffffu, exceed && valid);
        if (lane == 0) flag_buf[buf * 4 + wq] = w_any ? 1 : 0;
        lp_bar(1);
        const bool update = (flag_buf[buf * 4] | flag_buf[buf * 4 + 1] | flag_buf[buf * 4 + 2] | flag_buf[buf * 4 + 3]) != 0;
        if (update) {
          // move the references to the running maxima, rescale the sums and the context accumulators
#pragma unroll
          for (int h = 0; h < LP_NH; ++h)
            if (h < Hc) {
              const float mt = warp_max(valid ? __uint_as_float(sv[h]) : -INFINITY);
              if (lane == 0) red[wq * 32 + h] = mt;
            }
          lp_bar(2);
          if (tid < Hc) {
            const float mt = fmaxf(fmaxf(red[tid], red[32 + tid]), fmaxf(red[64 + tid], red[96 + tid]));
            const float m_old = m_buf[tid];
            const float m_new = fmaxf(m_old, mt);
            al_buf[tid] = ex2_approx((m_old - m_new) * sl2);     // 0 on the first tile
            m_buf[tid] = m_new;
          }
          lp_bar(3);
#pragma unroll
          for (int h = 0; h < LP_NH; ++h)
            if (h < Hc) {
              msl[h] = m_buf[h] * sl2;
              t[h] = fmaf(__uint_as_float(sv[h]), sl2, -msl[h]);
              l_part[h] *= al_buf[h];
            }
          if (j > sg.t_lo) {
            mbar_wait(&c_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
            tc_fence_after();
            for (int a = 0; a < NS; ++a) {
              uint32_t cv[32];
              tmem_ld_32x32(acc + a * LP_NH, cv);
              tmem_ld_wait();
#pragma unroll
              for (int h = 0; h < LP_NH; ++h)
                if (h < Hc) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * al_buf[h]);
              tmem_st_32x32(acc + a * LP_NH, cv);
            }
            tmem_st_wait();
          }
        }
        LP_T(t_mid);
        if (it >= 2) mbar_wait(&c_done[buf], ((it - 2) >> 1) & 1);     // P^T[buf] is no longer read by tile it - 2
        LP_T(t_pw);
        if (lane_on) {
          uint8_t* prow = pt + buf * LP_PT + p_off;
          if (valid) {
#pragma unroll
            for (int h = 0; h < LP_NH; ++h)
              if (h < Hc) {
                const __nv_bfloat16 pb = __float2bfloat16_rn(ex2_approx(t[h]));
                l_part[h] += __bfloat162float(pb);                     // the sums the tensor core will see
                *reinterpret_cast<__nv_bfloat16*>(prow + h * 128 + ((p_unit ^ (h & 7)) << 4)) = pb;
              }
          } else {
#pragma unroll
            for (int h = 0; h < LP_NH; ++h)
              if (h < Hc) *reinterpret_cast<uint16_t*>(prow + h * 128 + ((p_unit ^ (h & 7)) << 4)) = 0;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_ready[buf]);
        LP_T(t_p);
        // m_buf / al_buf / flag_buf[buf] are rewritten two tiles later at the earliest, behind lp_bar(1) of the next tile
      }
      // ---- end of the segment: row sums, then C^T / l -> ctx[part][clip]
#ifdef LP_TIMING
      const long long te0 = clock64();
#endif
#pragma unroll
      for (int h = 0; h < LP_NH; ++h)
        if (h < Hc) {
          const float v = warp_sum(l_part[h]);
          if (lane == 0) red[wq * 32 + h] = v;
        }
      lp_bar(2);
      const int part = sg.t_lo > 0 ? 1 : 0;
      if (tid < Hc) {
        const float l = (red[tid] + red[32 + tid]) + (red[64 + tid] + red[96 + tid]);
        linv_buf[tid] = 1.0f / l;
        if (ml != nullptr && rank == 0) {
          ml[(static_cast<long long>(part) * B + sg.clip) * 32 + tid] = make_float2(m_buf[tid] * sl2, l);
          if (sg.t_lo == 0 && sg.t_hi == n_tiles)   // the whole clip in one segment: no second part
            ml[(static_cast<long long>(B) + sg.clip) * 32 + tid] = make_float2(-INFINITY, 0.f);
        }
      }
      lp_bar(3);
      // thread = latent column; the last context MMAs of the segment have completed
      mbar_wait(&c_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* out0 = ctx + part * part_stride + static_cast<long long>(sg.clip) * H * d + col0 + tid;
      for (int a = 0; a < NS; ++a) {
        uint32_t cv[32];
        tmem_ld_32x32(acc + a * LP_NH, cv);
        tmem_ld_wait();
        __nv_bfloat16* out = out0 + a * 128;
#pragma unroll
        for (int h = 0; h < LP_NH; ++h)
          if (h < Hc) out[static_cast<long long>(h) * d] = __float2bfloat16_rn(__uint_as_float(cv[h]) * linv_buf[h]);
      }
      tc_fence_before();
      mbar_arrive(&c_free[seg & 1]);
      ++seg;
#ifdef LP_TIMING
      t_epi += clock64() - te0;
#endif
    }
#ifdef LP_TIMING
    if (blockIdx.x == 0 && tid == 0)
      printf("softmax per tile: wait s_full %lld | ld + send %lld | wait peer %lld | add..rescale %lld | wait c_done %lld | "
             "exp + P + arrive %lld ; epilogue per segment %lld (%d segments); total %lld clk for %d tiles\n", t_sf / it,
             t_st / it, t_xw / it, t_mid / it, t_pw / it, t_p / it, t_epi / seg, seg, clock64() - tt0, it);
#endif
  }

  // nobody leaves while the peer may still write into this CTA's exchange buffer or arrive on its barriers
  tc_fence_before();
  __syncthreads();
  lp_cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<LP_TMEM_COLS>(tmem_base);
  }
}

// Returns WF_ERR_UNSUPPORTED (without setting an error) when the shape does not fit the pair kernel: the caller falls
// back to the two-pass kernel of latent.cu.
bool latent_pair_supported(int H) {
  const int d = H * 64;
  if (d % 256 != 0) return false;
  const int hp = (H + 7) / 8 * 8, ns = d / 256;
  if (ns < LP_NI) return false;     // every issuer of pass A needs a stage in every tile
  if (LP_TMEM_S + 2 * ns * LP_NH > LP_TMEM_COLS) return false;
  const int fixed = 1024 + 2 * ns * hp * 128 + 2 * LP_PT + 2 * LP_XCH + LP_MISC;
  return ((LP_SMEM_LIMIT - fixed) / LP_STAGE & ~1) >= 2 * ns;
}

int latent_attention_pair(const void* qp, const void* src, void* ctx, float* ml, long long part_stride, int B, int T,
                          int H, cudaStream_t stream) {
  if (!latent_pair_supported(H)) return WF_ERR_UNSUPPORTED;
  const int d = H * 64;
  const int hp = (H + 7) / 8 * 8, ns = d / 256;
  const int fixed = 1024 + 2 * ns * hp * 128 + 2 * LP_PT + 2 * LP_XCH + LP_MISC;
  int nslot = (LP_SMEM_LIMIT - fixed) / LP_STAGE;
  if (nslot > LP_MAX_S) nslot = LP_MAX_S;
  nslot &= ~1;    // two issuers per pass deal the slots by parity: a slot is always consumed by the same thread
  const int smem = fixed + nslot * LP_STAGE;
  CUtensorMap mx, mq;
  int rc = make_map_bf16(&mx, src, static_cast<long long>(B) * T, d, d, LP_KT);
  if (rc) return rc;
  rc = make_map_bf16(&mq, qp, static_cast<long long>(B) * H, d, d, hp);
  if (rc) return rc;
  using Kern = void (*)(CUtensorMap, CUtensorMap, __nv_bfloat16*, float2*, long long, int, int, int, int, int, int, float);
  Kern kern = latent_pair_kernel<0>;
  int which = 0;
  switch (H) {   // Whisper widths: small 12, medium 16, large 20 heads
    case 12: kern = latent_pair_kernel<12>; which = 1; break;
    case 16: kern = latent_pair_kernel<16>; which = 2; break;
    case 20: kern = latent_pair_kernel<20>; which = 3; break;
    default: break;
  }
  static PerDeviceOnce configured[4];  // function attributes are per device
  if (configured[which].first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LP_SMEM_LIMIT));
  }
  const float sl2 = 0.125f * 1.44269504088896340736f;   // 64^-0.5 * log2(e)
  // one cluster per clip while they all fit at once (or when the caller has no room for a second part); otherwise one
  // cluster per SM pair, each with 1 / clusters of the tile sequence (>= one clip: a clip is cut at most once)
  int clusters = B;
  if (ml != nullptr && B > num_sms() / 2) clusters = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(LP_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = 2;
  attr[n].val.clusterDim.y = 1;
  attr[n].val.clusterDim.z = 1;
  ++n;
  if (pdl_enabled(2)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  const int prio = launch_priority(2);
  if (prio != INT_MIN) {
    attr[n].id = cudaLaunchAttributePriority;
    attr[n].val.priority = prio;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  WF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mq, reinterpret_cast<__nv_bfloat16*>(ctx),
                                   reinterpret_cast<float2*>(ml), part_stride, B, T, H, hp, ns, nslot, sl2));
  count_launch();
  return WF_OK;
}

}  // namespace wf
