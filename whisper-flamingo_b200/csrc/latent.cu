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
// latent_attn_kernel: one CLUSTER of CS = d / 256 CTAs per clip; CTA r owns latent columns [256 r, 256 r + 256) of
// src and of q', and the softmax bookkeeping of heads [4 r, 4 r + 4).  Per 128-key tile j:
//   TMA      : src tile slice [128 keys x 256] -> 4 swizzled 16 KB chunks (2-stage ring), read from HBM exactly once;
//   tcgen05  : S^T_partial[128 keys x 32 heads] = tile (A, K-major) x q'_slice^T (B)           -> TMEM
//   softmax warps (thread = key): push the partial scores of 4 heads to their owner CTA (st.shared::cluster), owner sums
//              the CS partials, keeps the running max / sum of its 4 heads, pushes p (bf16) and the rescale factor
//              alpha of its heads to every CTA of the cluster;
//   tcgen05  : C^T[256 latent x 32 heads] += tile^T (A, MN-major: the SAME smem chunks) x P^T (B)   in TMEM
// so each CTA ends with its 256-column slice of c_h for every head and writes it normalised by 1 / l_h.
#include "common.cuh"
#include "kernels.h"

namespace wf {

static constexpr int LA_KEYS = 128;                 // keys per tile
static constexpr int LA_DS = 256;                   // latent columns per CTA
static constexpr int LA_CH = LA_DS / 64;            // 64-column swizzle atoms (chunks) per tile slice
static constexpr int LA_NH = 32;                    // head columns of both MMAs (H <= 32; unused ones are ignored)
static constexpr int LA_HPC = LA_DS / 64;           // heads whose softmax a CTA owns (head_dim 64 => H = 4 CS)
static constexpr int LA_MAX_CS = 5;                 // d <= 1280
static constexpr int LA_CHUNK = LA_KEYS * 128;      // 16 KB
static constexpr int LA_STAGE = LA_CH * LA_CHUNK;   // 64 KB
static constexpr int LA_STAGES = 2;
static constexpr int LA_QCH = LA_NH * 128;          // 4 KB: one K-major atom of q' (and of P^T)
static constexpr int LA_OFF_Q = LA_STAGES * LA_STAGE;
static constexpr int LA_OFF_PO = LA_OFF_Q + LA_CH * LA_QCH;
static constexpr int LA_PO = 2 * LA_QCH;            // P^T operand [32 heads x 128 keys] = two atoms of 64 keys
static constexpr int LA_OFF_PI = LA_OFF_PO + 2 * LA_PO;
static constexpr int LA_PI = LA_NH * LA_KEYS * 2;   // P^T landing buffer (plain [head][key] bf16), filled by the owners
static constexpr int LA_OFF_XS = LA_OFF_PI + 2 * LA_PI;
static constexpr int LA_XS = LA_MAX_CS * LA_HPC * LA_KEYS * 4;  // partial scores [source CTA][own head][key] fp32
static constexpr int LA_OFF_AL = LA_OFF_XS + 2 * LA_XS;         // alpha [2][32] | 1/l [32] | red [3][4 warps][4] fp32
static constexpr int LA_OFF_BAR = LA_OFF_AL + 1024;
static constexpr int LA_SMEM = LA_OFF_BAR + 256 + 1024;
static constexpr int LA_TMEM_COLS = 128;            // S^T(0) S^T(1) C^T(rows 0..127) C^T(rows 128..255), 32 each
static constexpr int LA_THREADS = 256;

__device__ __forceinline__ void la_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t la_mapa(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void la_st_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void la_st_u16(uint32_t cluster_addr, uint16_t v) {
  asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(cluster_addr), "h"(v) : "memory");
}
__device__ __forceinline__ void la_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void la_fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void la_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool la_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait for arrivals made by OTHER CTAs of the cluster (their st.shared::cluster data must be visible afterwards)
__device__ __forceinline__ void la_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (la_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!la_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("libwf: cluster mbarrier wait timeout (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void la_st_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void la_st_v2(uint32_t cluster_addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(a), "r"(b) : "memory");
}

// -DLA_TIMING: thread 128 of block (0, 0) accumulates the cycles of each phase of its per-tile chain and prints them
#ifdef LA_TIMING
#define LA_TICK0() long long la_t = clock64()
#define LA_TICKW() do { const long long n_ = clock64(); la_acc[6] += n_ - la_t; la_t = n_; } while (0)
#define LA_TICK(i) do { const long long n_ = clock64(); la_acc[i] += n_ - la_t; la_t = n_; } while (0)
#define LA_REPORT()                                                                                                 \
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)                                                               \
    printf("latent chain cycles/tile: wait_s %lld | ld+push_xs %lld | wait_xs %lld | softmax+push_p %lld | wait_pb " \
           "%lld | stage+rescale %lld | (ld..tick0 %lld) tiles %d\n", la_acc[6] / n_tiles, la_acc[1] / n_tiles,      \
           la_acc[2] / n_tiles, la_acc[3] / n_tiles, la_acc[4] / n_tiles, la_acc[5] / n_tiles, la_acc[0] / n_tiles,  \
           n_tiles)
#else
#define LA_TICK0()
#define LA_TICKW()
#define LA_TICK(i)
#define LA_REPORT()
#endif

// bf16 x bf16 -> fp32, A MN-major (bit 15), B K-major
__host__ __device__ constexpr uint32_t la_idesc_a_mn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 15); }

__global__ void __launch_bounds__(LA_THREADS, 1)
latent_attn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q,
                   __nv_bfloat16* __restrict__ ctx, int T, int H, int CS, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared window starts at the same offset in every CTA of the cluster, so the aligned base does too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LA_OFF_BAR);
  uint64_t* q_full = bars;             // [1]
  uint64_t* st_full = bars + 1;        // [2]  tile slice landed
  uint64_t* st_empty = bars + 3;       // [2]  C^T MMAs of the tile have read it
  uint64_t* s_full = bars + 5;         // [2]  S^T_partial(j) in TMEM
  uint64_t* s_free = bars + 7;         // [2]  ... read into registers (128 arrivals)
  uint64_t* xs_full = bars + 9;        // [2]  partial scores of my heads arrived from all CS CTAs
  uint64_t* pb_full = bars + 11;       // [2]  P^T rows + alpha of all heads arrived from all CS owners
  uint64_t* p_ready = bars + 13;       // [2]  P^T operand staged, C^T rescaled (128 arrivals)
  uint64_t* c_done = bars + 15;        // [2]  C^T MMAs of the tile completed
  uint64_t* fin_full = bars + 17;      // [1]  1 / l of all heads arrived
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x;            // rank in the cluster (cluster = the CS CTAs along x) = latent slice
  const int b = blockIdx.y;            // clip
  const int n_tiles = (T + LA_KEYS - 1) / LA_KEYS;
  const int d = CS * LA_DS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_q);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128);
      mbar_init(&xs_full[i], CS); mbar_init(&pb_full[i], CS); mbar_init(&p_ready[i], 128); mbar_init(&c_done[i], 1);
    }
    mbar_init(fin_full, CS);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<LA_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  la_cluster_sync();   // every CTA's barriers exist before anyone arrives on them remotely
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_arrive_expect_tx(q_full, LA_CH * LA_QCH);
    for (int c = 0; c < LA_CH; ++c)
      tma_load_2d(smem + LA_OFF_Q + c * LA_QCH, &map_q, q_full, r * LA_DS + c * 64, b * H);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      mbar_wait(&st_empty[st], ((j >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&st_full[st], LA_STAGE);
      for (int c = 0; c < LA_CH; ++c)
        tma_load_2d(smem + st * LA_STAGE + c * LA_CHUNK, &map_x, &st_full[st], r * LA_DS + c * 64, b * T + j * LA_KEYS);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(LA_KEYS, LA_NH);    // S^T = tile q'^T : both K-major
    constexpr uint32_t idesc_c = la_idesc_a_mn(128, LA_NH);          // C^T += tile^T P^T : A MN-major
    auto issue_s = [&](int j) {
      const int st = j & 1;
#pragma unroll
      for (int c = 0; c < LA_CH; ++c) {
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem + st * LA_STAGE + c * LA_CHUNK));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem + LA_OFF_Q + c * LA_QCH));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + st * LA_NH, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (c | k) != 0);
      }
      umma_commit(&s_full[st]);
    };
    auto issue_c = [&](int j) {
      const int st = j & 1;
      const uint32_t po = smem_u32(smem + LA_OFF_PO + st * LA_PO);
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const uint32_t a_addr = smem_u32(smem + st * LA_STAGE + 2 * mb * LA_CHUNK);
#pragma unroll
        for (int kk = 0; kk < LA_KEYS / 16; ++kk) {
          // A: 16 keys (two 8-row groups of 1024 B) x 128 latent columns (two 64-column atoms, LBO = one chunk)
          const uint64_t a_desc = umma_desc_mnmajor_sw128(a_addr + kk * 2048);
          const uint64_t b_desc = umma_desc_kmajor_sw128(po + (kk >> 2) * LA_QCH) + 2 * (kk & 3);
          umma_f16(tmem_base + 2 * LA_NH + mb * LA_NH, a_desc, b_desc, idesc_c, (j > 0 || kk > 0) ? 1u : 0u);
        }
      }
      umma_commit(&st_empty[st]);
      umma_commit(&c_done[st]);
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j <= n_tiles; ++j) {
      bool need_s = j < n_tiles, need_c = j > 0;
      const long long t0 = clock64();
      while (need_s || need_c) {
        if (need_c && mbar_try_wait(&p_ready[(j - 1) & 1], ((j - 1) >> 1) & 1)) {
          tc_fence_after();
          issue_c(j - 1);
          need_c = false;
        }
        if (need_s && mbar_try_wait(&st_full[j & 1], (j >> 1) & 1) &&
            (j < 2 || mbar_try_wait(&s_free[j & 1], ((j >> 1) - 1) & 1))) {
          tc_fence_after();
          issue_s(j);
          need_s = false;
        }
        if (clock64() - t0 > 4000000000LL) {
          printf("libwf: latent attention MMA issuer timeout (block %d,%d tile %d need_s %d need_c %d)\n", blockIdx.x,
                 blockIdx.y, j, int(need_s), int(need_c));
          __trap();
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax (thread = key = TMEM lane)
    const int tid = threadIdx.x - 128;
    const int qd = warp - 4;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    uint32_t rb[8];                                                   // my smem base as seen in CTA c's window
#pragma unroll
    for (int c = 0; c < 8; ++c) rb[c] = la_mapa(smem_u32(smem), c < CS ? c : 0);
    float* al_buf = reinterpret_cast<float*>(smem + LA_OFF_AL);       // [2][32]
    float* linv_buf = al_buf + 64;                                    // [32]
    float* red = al_buf + 96;                                         // [3][4][4]
    float m_run[LA_HPC], l_part[LA_HPC];
#ifdef LA_TIMING
    long long la_acc[7] = {0, 0, 0, 0, 0, 0, 0};
#endif
#pragma unroll
    for (int i = 0; i < LA_HPC; ++i) { m_run[i] = -INFINITY; l_part[i] = 0.f; }

    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      LA_TICK0();
      mbar_wait(&s_full[buf], ph);
      tc_fence_after();
      LA_TICKW();
      uint32_t sv[32];
      tmem_ld_32x32(lane_addr + buf * LA_NH, sv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[buf]);
      LA_TICK(0);
      // ---- partial scores of heads 4c .. 4c+3 -> their owner CTA c, slot [my rank][key][4] (one 16-byte store each)
      {
        const uint32_t xoff = LA_OFF_XS + buf * LA_XS + (r * LA_KEYS + tid) * 16;
#pragma unroll
        for (int c = 0; c < LA_MAX_CS; ++c)
          if (c < CS) la_st_v4(rb[c] + xoff, sv[4 * c], sv[4 * c + 1], sv[4 * c + 2], sv[4 * c + 3]);
      }
      la_fence_cluster();
      la_bar(1);
      if (tid == 0)
        for (int c = 0; c < CS; ++c) la_arrive_remote(la_mapa(smem_u32(&xs_full[buf]), c));
      LA_TICK(1);
      la_wait_cluster(&xs_full[buf], ph);
      LA_TICK(2);
      // ---- owner: full scores of my 4 heads for this key, running max over the tile's 128 keys
      const float4* xs = reinterpret_cast<const float4*>(smem + LA_OFF_XS + buf * LA_XS);
      const bool valid = j * LA_KEYS + tid < T;
      float s[LA_HPC] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < CS; ++c) {
        const float4 v = xs[c * LA_KEYS + tid];
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      }
#pragma unroll
      for (int hh = 0; hh < LA_HPC; ++hh) {
        const float mt = warp_max(valid ? s[hh] : -INFINITY);
        if (lane == 0) red[(buf * 4 + qd) * 4 + hh] = mt;
      }
      la_bar(2);
      float alpha[LA_HPC], p[LA_HPC];
#pragma unroll
      for (int hh = 0; hh < LA_HPC; ++hh) {
        float mt = fmaxf(fmaxf(red[(buf * 4 + 0) * 4 + hh], red[(buf * 4 + 1) * 4 + hh]),
                         fmaxf(red[(buf * 4 + 2) * 4 + hh], red[(buf * 4 + 3) * 4 + hh]));
        const float m_new = fmaxf(m_run[hh], mt);
        alpha[hh] = ex2_approx((m_run[hh] - m_new) * sl2);            // 0 on the first tile
        p[hh] = valid ? ex2_approx((s[hh] - m_new) * sl2) : 0.f;
        m_run[hh] = m_new;
      }
      const uint32_t p01 = pack_bf16(p[0], p[1]), p23 = pack_bf16(p[2], p[3]);
      l_part[0] = l_part[0] * alpha[0] + bf16lo(p01);                 // the sums the tensor core will see
      l_part[1] = l_part[1] * alpha[1] + bf16hi(p01);
      l_part[2] = l_part[2] * alpha[2] + bf16lo(p23);
      l_part[3] = l_part[3] * alpha[3] + bf16hi(p23);
      // ---- push p (bf16, [owner][key][4 heads]: one 8-byte store) and alpha of my heads to every CTA of the cluster
      {
        const uint32_t poff = LA_OFF_PI + buf * LA_PI + (r * LA_KEYS + tid) * 8;
#pragma unroll
        for (int c = 0; c < LA_MAX_CS; ++c)
          if (c < CS) la_st_v2(rb[c] + poff, p01, p23);
        if (tid < LA_HPC * CS) {
          const int hh = tid & 3, c = tid >> 2;
          const float a = hh == 0 ? alpha[0] : hh == 1 ? alpha[1] : hh == 2 ? alpha[2] : alpha[3];
          la_st_f32(la_mapa(smem_u32(smem + LA_OFF_AL + (buf * 32 + r * LA_HPC + hh) * 4), c), a);
        }
      }
      la_fence_cluster();
      la_bar(1);
      if (tid == 0)
        for (int c = 0; c < CS; ++c) la_arrive_remote(la_mapa(smem_u32(&pb_full[buf]), c));
      LA_TICK(3);
      la_wait_cluster(&pb_full[buf], ph);
      LA_TICK(4);
      // ---- stage P^T as the K-major swizzled B operand [32 heads x 128 keys]: unit = (owner c, 8 keys) = 64 bytes in,
      //      four 16-byte rows (one per head) out
      if (tid < CS * 16) {
        const int c = tid >> 4, g = tid & 15;
        const uint4* pi = reinterpret_cast<const uint4*>(smem + LA_OFF_PI + buf * LA_PI + (c * LA_KEYS + g * 8) * 8);
        const uint4 v0 = pi[0], v1 = pi[1], v2 = pi[2], v3 = pi[3];     // key k: (h0 h1 | h2 h3), two keys per uint4
        uint8_t* po = smem + LA_OFF_PO + buf * LA_PO + (g >> 3) * LA_QCH;
        const uint32_t lo[8] = {v0.x, v0.z, v1.x, v1.z, v2.x, v2.z, v3.x, v3.z};
        const uint32_t hi[8] = {v0.y, v0.w, v1.y, v1.w, v2.y, v2.w, v3.y, v3.w};
#pragma unroll
        for (int hh = 0; hh < LA_HPC; ++hh) {
          const int h = c * LA_HPC + hh;
          const uint32_t sel = (hh & 1) ? 0x7632u : 0x5410u;
          uint4 o;
          if (hh < 2) {
            o.x = __byte_perm(lo[0], lo[1], sel); o.y = __byte_perm(lo[2], lo[3], sel);
            o.z = __byte_perm(lo[4], lo[5], sel); o.w = __byte_perm(lo[6], lo[7], sel);
          } else {
            o.x = __byte_perm(hi[0], hi[1], sel); o.y = __byte_perm(hi[2], hi[3], sel);
            o.z = __byte_perm(hi[4], hi[5], sel); o.w = __byte_perm(hi[6], hi[7], sel);
          }
          *reinterpret_cast<uint4*>(po + h * 128 + (((g & 7) ^ (h & 7)) << 4)) = o;
        }
      }
      // ---- C^T *= alpha (column h), only when some head moved its maximum; C^T MMAs of tile j-1 must be complete
      if (j > 0) {
        bool any = false;
        for (int h = 0; h < H; ++h) any = any || (al_buf[buf * 32 + h] != 1.0f);
        mbar_wait(&c_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (any) {
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            uint32_t cv[32];
            tmem_ld_32x32(lane_addr + 2 * LA_NH + mb * LA_NH, cv);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < LA_NH; ++h)
              if (h < H) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * al_buf[buf * 32 + h]);
            tmem_st_32x32(lane_addr + 2 * LA_NH + mb * LA_NH, cv);
          }
          tmem_st_wait();
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
      LA_TICK(5);
    }
    LA_REPORT();
    // ---- 1 / l of my heads -> every CTA
    {
      float l[LA_HPC];
#pragma unroll
      for (int hh = 0; hh < LA_HPC; ++hh) {
        const float v = warp_sum(l_part[hh]);
        if (lane == 0) red[(2 * 4 + qd) * 4 + hh] = v;
      }
      la_bar(2);
#pragma unroll
      for (int hh = 0; hh < LA_HPC; ++hh)
        l[hh] = (red[(2 * 4 + 0) * 4 + hh] + red[(2 * 4 + 1) * 4 + hh]) + (red[(2 * 4 + 2) * 4 + hh] + red[(2 * 4 + 3) * 4 + hh]);
      if (tid < LA_HPC * CS) {
        const int hh = tid & 3, c = tid >> 2;
        const float v = hh == 0 ? l[0] : hh == 1 ? l[1] : hh == 2 ? l[2] : l[3];
        la_st_f32(la_mapa(smem_u32(smem + LA_OFF_AL + (64 + r * LA_HPC + hh) * 4), c), 1.0f / v);
      }
      la_fence_cluster();
      la_bar(1);
      if (tid == 0)
        for (int c = 0; c < CS; ++c) la_arrive_remote(la_mapa(smem_u32(fin_full), c));
      la_wait_cluster(fin_full, 0);
    }
    // ---- epilogue: c_h[256 r + 128 mb + tid] = C^T[row][h] / l_h  (thread = latent column, coalesced over the warp)
    mbar_wait(&c_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      uint32_t cv[32];
      tmem_ld_32x32(lane_addr + 2 * LA_NH + mb * LA_NH, cv);
      tmem_ld_wait();
      __nv_bfloat16* out = ctx + static_cast<long long>(b) * H * d + r * LA_DS + mb * 128 + tid;
#pragma unroll
      for (int h = 0; h < LA_NH; ++h)
        if (h < H) out[static_cast<long long>(h) * d] = __float2bfloat16_rn(__uint_as_float(cv[h]) * linv_buf[h]);
    }
  }

  tc_fence_before();
  __syncthreads();
  la_cluster_sync();   // no CTA leaves (and frees its shared memory) while a peer may still store into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<LA_TMEM_COLS>(tmem_base);
  }
}

int latent_attention(const void* qp, const void* src, void* ctx, int B, int T, int H, cudaStream_t stream) {
  const int d = H * 64;
  WF_REQUIRE(B > 0 && T > 0 && H > 0 && d % LA_DS == 0 && d / LA_DS <= LA_MAX_CS,
             "latent attention: needs head_dim 64 and n_state in {256, 512, 768, 1024, 1280} (got %d heads)", H);
  const int cs = d / LA_DS;
  CUtensorMap mx, mq;
  int rc = make_map_bf16(&mx, src, static_cast<long long>(B) * T, d, d, LA_KEYS);
  if (rc) return rc;
  rc = make_map_bf16(&mq, qp, static_cast<long long>(B) * H, d, d, LA_NH);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LA_SMEM));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs, B);
  cfg.blockDim = dim3(LA_THREADS);
  cfg.dynamicSmemBytes = LA_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float sl2 = 0.125f * 1.44269504088896340736f;   // 64^-0.5 * log2(e)
  WF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, latent_attn_kernel, mx, mq, reinterpret_cast<__nv_bfloat16*>(ctx), T, H, cs,
                                   sl2));
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
