// NOT PART OF THE LIBRARY - kept as the record of a measured dead end (profiles/r02_latent_experiments.txt, section 2).
// One-pass latent attention with the whole 64-key tile (160 KB at d = 1280) resident in ONE SM: the ring holds a single
// tile, so load latency, both products and the softmax hand-off of every tile run in series: 210-218 us against 114 us
// for the two-pass kernel and 102.7 us for the 2-CTA-cluster kernel that replaced it (csrc/latent_pair.cu).  It was
// parity-green (tests/test_kernels_gpu.py -k latent) when it was wired into csrc/latent.cu.
//
// One-token cross-attention over the source rows, ONE pass: the 64-key tile stays in shared memory between the score
// product and the context product.  Same contract as latent_attn_kernel (latent.cu: q' = Wk_h^T q_h, context
// c_h = softmax(src q'_h / 8)^T src); replaces the per-step recompute of reference whisper/decoding.py:155-164 for the
// cross-attention (model.py:93-108 with xa) and the gated x-attention (model.py:110-134 with xt).
//
// Why a second kernel: the two-pass kernel reads every 128-key tile twice (HBM, then L2).  On B200 the L2 slices
// deliver ~6300 B/clk chip-wide (~12 TB/s, only 2x HBM), so its 2 x 491 MB (+ q' re-reads) per launch cost >= 86 us
// of L2 time before any latency: it ran at 115 us = 61 % of the HBM roofline and could not go below ~90.  Here a tile
// is 64 keys x d (160 KB at d = 1280) and lives in a ring of 16 KB stages [64 keys x 128 columns] that BOTH products
// read from shared memory; a stage is released by the commit of its context MMAs and refilled with the next tile:
//   pass A  S^T[64 keys x 32 heads] += stage (A, K-major, two swizzle atoms) x q'^T (B; q' is resident: 60 KB at d = 1280,
//           which leaves the ring exactly one tile - streaming q' through a 3-slot ring of its own was measured first:
//           210 us, every stage waited ~1 us for its q' atoms from L2)
//   softmax thread = key (a 64-row accumulator keeps rows 16 q .. 16 q + 15 in lanes 32 q .. 32 q + 15); reference
//           maximum with a 2^8 window as in latent.cu; P^T (bf16) -> smem
//   pass B  C^T[128 columns x 32 heads] += stage^T (A, MN-major: the same swizzled bytes) x P^T (B), d / 128
//           accumulators in TMEM; its commit frees the stage
// The ring holds one tile (plus what is left of shared memory), so the products of consecutive tiles do not overlap:
// stage a of tile j + 1 is requested when the context MMAs of stage a of tile j retire, the scores of tile j + 1 follow
// the arrivals.  L2 -> SM traffic per launch: the source, once.  MMA count per 64 keys: 80 + 40 narrow tcgen05.mma
// (40 clk each on the SM's tensor front end): 58 us per 1500-key clip, under the HBM time.
#include "common.cuh"
#include "kernels.h"

namespace wf {

static constexpr int L1_KT = 64;                    // keys per tile
static constexpr int L1_BOX = L1_KT * 128;          // 8 KB: [64 keys x 64 columns] bf16, 128B-swizzled = one TMA box
static constexpr int L1_STAGE = 2 * L1_BOX;         // 16 KB: 128 columns = the A operand of one context accumulator
static constexpr int L1_NH = 32;                    // head columns of both MMAs (H <= 32)
static constexpr int L1_PT = L1_NH * 128;           // 4 KB: P^T rows (heads) x 64 keys
static constexpr int L1_MISC = 1280;
static constexpr int L1_MAX_S = 16;
static constexpr int L1_SMEM_LIMIT = 227 * 1024;
static constexpr int L1_TMEM_COLS = 512;            // S^T (2 x 32) | C^T (d / 128 accumulators x 32)
static constexpr int L1_TMEM_C = 2 * L1_NH;
static constexpr int L1_THREADS = 256;

__device__ __forceinline__ void l1_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void l1_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// MN-major operand, 128B swizzle: rows = K index (keys; 128 B = 64 MN elements each), 8-row groups 1024 B apart (SBO),
// the next 64 MN elements one TMA box further (LBO)
__device__ __forceinline__ uint64_t l1_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(L1_BOX >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t l1_idesc_a_mn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 15); }

// grid = clips; NS = d / 128 stages per tile; NSLOT ring stages (>= NS)
__global__ void __launch_bounds__(L1_THREADS, 1)
latent_attn1_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q,
                    __nv_bfloat16* __restrict__ ctx, int T, int H, int HP, int NS, int NSLOT, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_atom = HP * 128;
  uint8_t* ring = smem;                              // NSLOT stages
  uint8_t* qs = ring + NSLOT * L1_STAGE;             // q': d / 64 K-major atoms of HP rows (heads)
  uint8_t* pt = qs + 2 * NS * q_atom;                // P^T operand
  uint8_t* misc = pt + L1_PT;
  float* m_buf = reinterpret_cast<float*>(misc);     // [32] reference maximum of each head
  float* al_buf = m_buf + 32;                        // [32] rescale factor when the reference moved
  float* linv_buf = al_buf + 32;                     // [32]
  float* red = linv_buf + 32;                        // [4][32]
  int* flag_buf = reinterpret_cast<int*>(red + 128); // [2][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 928);
  uint64_t* full_s = bars;                           // [NSLOT] stage landed
  uint64_t* empty_s = full_s + L1_MAX_S;             // [NSLOT] context MMAs of the stage completed
  uint64_t* q_full = empty_s + L1_MAX_S;
  uint64_t* s_full = q_full + 1;                     // [2] scores of a tile in TMEM
  uint64_t* s_free = s_full + 2;                     // [2] ... copied to registers (128 arrivals)
  uint64_t* p_ready = s_free + 2;                    // P^T staged, C^T rescaled (128 arrivals)
  uint64_t* c_done = p_ready + 1;                    // context MMAs of a tile completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;                          // clip
  const int n_tiles = (T + L1_KT - 1) / L1_KT;
  const int d = NS * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_q);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full_s[i], 1); mbar_init(&empty_s[i], 1); }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128); }
    mbar_init(p_ready, 128);
    mbar_init(c_done, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<L1_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer: source stages (static data: the
    // first NSLOT stages are requested before the previous kernel has finished)
    const uint64_t once = l2_policy_evict_first();
    int slot = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int row = b * T + j * L1_KT;
      for (int a = 0; a < NS; ++a) {
        mbar_wait(&empty_s[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_s[slot], L1_STAGE);
        uint8_t* dst = ring + slot * L1_STAGE;
        l1_tma_load_2d(dst, &map_x, &full_s[slot], a * 128, row, once);
        l1_tma_load_2d(dst + L1_BOX, &map_x, &full_s[slot], a * 128 + 64, row, once);
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2 && lane == 0) {
    // ------------------------------------------------------------------ q' (resident for the whole kernel)
    pdl_wait();       // q' comes from the previous kernel
    mbar_arrive_expect_tx(q_full, 2 * NS * q_atom);
    for (int i = 0; i < 2 * NS; ++i) tma_load_2d(qs + i * q_atom, &map_q, q_full, i * 64, b * H);
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer, pass A
    // S^T[64 keys x 32 heads] = stage (A, K-major) x q'^T (B, K-major; rows >= HP of an atom alias what follows it and
    // only produce head columns nobody reads), K = d in 16-column steps
    constexpr uint32_t idesc_s = umma_idesc_bf16(L1_KT, L1_NH);
    const uint32_t ra = smem_u32(ring), qa = smem_u32(qs);
    int slot = 0;
    uint32_t phase = 0;
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j >= 2) mbar_wait(&s_free[j & 1], ((j >> 1) - 1) & 1);
      tc_fence_after();
      for (int a = 0; a < NS; ++a) {
        mbar_wait(&full_s[slot], phase);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(ra + slot * L1_STAGE + half * L1_BOX);
          const uint64_t b_desc = umma_desc_kmajor_sw128(qa + (2 * a + half) * q_atom);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + (j & 1) * L1_NH, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (a | half | k) != 0);
        }
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
      }
      umma_commit(&s_full[j & 1]);
    }
  } else if (warp == 3 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer, pass B
    // C^T[128 columns x 32 heads] (+)= stage^T (A, MN-major: rows = keys, two 64-column atoms 8 KB apart) x P^T (B,
    // K-major, one atom of 64 keys), K = 64 keys in 16-key steps.  The stage was waited for by pass A of the same tile.
    constexpr uint32_t idesc_c = l1_idesc_a_mn(128, L1_NH);
    const uint32_t ra = smem_u32(ring), pa = smem_u32(pt);
    int slot = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(p_ready, j & 1);
      tc_fence_after();
      const uint64_t p_desc = umma_desc_kmajor_sw128(pa);
      for (int a = 0; a < NS; ++a) {
        const uint32_t st = ra + slot * L1_STAGE;
#pragma unroll
        for (int kk = 0; kk < L1_KT / 16; ++kk)
          umma_f16(tmem_base + L1_TMEM_C + a * L1_NH, l1_desc_mn(st + kk * 2048), p_desc + 2 * kk, idesc_c,
                   (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty_s[slot]);
        if (++slot == NSLOT) slot = 0;
      }
      umma_commit(c_done);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax (thread = key = TMEM lane of S^T)
    const int wq = warp - 4;
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float l_part[L1_NH];
#pragma unroll
    for (int h = 0; h < L1_NH; ++h) l_part[h] = 0.f;
    if (tid < L1_NH) m_buf[tid] = -INFINITY;
    pdl_wait();     // the context rows are read by an earlier kernel of the stream
    l1_bar(1);
    // a 64-row accumulator keeps rows 16 q .. 16 q + 15 in lanes 32 q .. 32 q + 15 (profiles/r01_probe_tmem_m64_layout.txt)
    const int key = wq * 16 + (lane & 15);
    const bool lane_on = lane < 16;
    // P^T element (head h, my key): row = head (128 B), 16-byte units swizzled by the row
    const uint32_t p_off = (key & 7) * 2;
    const uint32_t p_unit = key >> 3;
    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[buf], ph);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(lane_base + buf * L1_NH, sv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[buf]);
      const bool valid = lane_on && j * L1_KT + key < T;
      // does any score leave the window of its head's reference maximum?  (always on the first tile)
      bool exceed = false;
#pragma unroll
      for (int h = 0; h < L1_NH; ++h)
        if (h < H) exceed = exceed || (__uint_as_float(sv[h]) - m_buf[h]) * sl2 > 8.f;
      const bool w_any = __any_sync(0xffffffffu, exceed && valid);
      if (lane == 0) flag_buf[buf * 4 + wq] = w_any ? 1 : 0;
      l1_bar(1);
      const bool update = (flag_buf[buf * 4] | flag_buf[buf * 4 + 1] | flag_buf[buf * 4 + 2] | flag_buf[buf * 4 + 3]) != 0;
      if (update) {
        // move the references to the running maxima, rescale the sums and the context accumulators
#pragma unroll
        for (int h = 0; h < L1_NH; ++h)
          if (h < H) {
            const float mt = warp_max(valid ? __uint_as_float(sv[h]) : -INFINITY);
            if (lane == 0) red[wq * 32 + h] = mt;
          }
        l1_bar(2);
        if (tid < H) {
          const float mt = fmaxf(fmaxf(red[tid], red[32 + tid]), fmaxf(red[64 + tid], red[96 + tid]));
          const float m_old = m_buf[tid];
          const float m_new = fmaxf(m_old, mt);
          al_buf[tid] = ex2_approx((m_old - m_new) * sl2);     // 0 on the first tile
          m_buf[tid] = m_new;
        }
        l1_bar(3);
#pragma unroll
        for (int h = 0; h < L1_NH; ++h)
          if (h < H) l_part[h] *= al_buf[h];
        if (j > 0) {
          mbar_wait(c_done, (j - 1) & 1);
          tc_fence_after();
          for (int a = 0; a < NS; ++a) {
            uint32_t cv[32];
            tmem_ld_32x32(lane_base + L1_TMEM_C + a * L1_NH, cv);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < L1_NH; ++h)
              if (h < H) cv[h] = __float_as_uint(__uint_as_float(cv[h]) * al_buf[h]);
            tmem_st_32x32(lane_base + L1_TMEM_C + a * L1_NH, cv);
          }
          tmem_st_wait();
        }
      }
      if (j >= 1) mbar_wait(c_done, (j - 1) & 1);     // P^T is no longer read by tile j - 1
      if (lane_on) {
        uint8_t* prow = pt + p_off;
#pragma unroll
        for (int h = 0; h < L1_NH; ++h)
          if (h < H) {
            const float p = valid ? ex2_approx((__uint_as_float(sv[h]) - m_buf[h]) * sl2) : 0.f;
            const __nv_bfloat16 pb = __float2bfloat16_rn(p);
            l_part[h] += __bfloat162float(pb);                       // the sums the tensor core will see
            *reinterpret_cast<__nv_bfloat16*>(prow + h * 128 + ((p_unit ^ (h & 7)) << 4)) = pb;
          }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // m_buf / al_buf / flag_buf[buf] are rewritten two tiles later at the earliest, behind l1_bar(1) of tile j + 1
    }
    // ---- 1 / l
#pragma unroll
    for (int h = 0; h < L1_NH; ++h)
      if (h < H) {
        const float v = warp_sum(l_part[h]);
        if (lane == 0) red[wq * 32 + h] = v;
      }
    l1_bar(2);
    if (tid < H) linv_buf[tid] = 1.0f / ((red[tid] + red[32 + tid]) + (red[64 + tid] + red[96 + tid]));
    l1_bar(3);
    // ---- epilogue: ctx[b, h, 128 a + tid] = C^T[a][tid][h] / l_h   (thread = latent column)
    mbar_wait(c_done, (n_tiles - 1) & 1);
    tc_fence_after();
    for (int a = 0; a < NS; ++a) {
      uint32_t cv[32];
      tmem_ld_32x32(lane_base + L1_TMEM_C + a * L1_NH, cv);
      tmem_ld_wait();
      __nv_bfloat16* out = ctx + static_cast<long long>(b) * H * d + a * 128 + tid;
#pragma unroll
      for (int h = 0; h < L1_NH; ++h)
        if (h < H) out[static_cast<long long>(h) * d] = __float2bfloat16_rn(__uint_as_float(cv[h]) * linv_buf[h]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<L1_TMEM_COLS>(tmem_base);
  }
}

// Returns WF_ERR_UNSUPPORTED (without setting an error) when the tile ring does not fit: the caller falls back to the
// two-pass kernel.
int latent_attention_onepass(const void* qp, const void* src, void* ctx, int B, int T, int H, cudaStream_t stream) {
  const int d = H * 64;
  const int hp = (H + 7) / 8 * 8, ns = d / 128;
  if (ns > (L1_TMEM_COLS - L1_TMEM_C) / L1_NH) return WF_ERR_UNSUPPORTED;
  const int fixed = 1024 + 2 * ns * hp * 128 + L1_PT + L1_MISC;
  int nslot = (L1_SMEM_LIMIT - fixed) / L1_STAGE;
  if (nslot > L1_MAX_S) nslot = L1_MAX_S;
  if (nslot < ns) return WF_ERR_UNSUPPORTED;
  const int smem = fixed + nslot * L1_STAGE;
  CUtensorMap mx, mq;
  int rc = make_map_bf16(&mx, src, static_cast<long long>(B) * T, d, d, L1_KT);
  if (rc) return rc;
  rc = make_map_bf16(&mq, qp, static_cast<long long>(B) * H, d, d, hp);
  if (rc) return rc;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(latent_attn1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L1_SMEM_LIMIT));
  }
  const float sl2 = 0.125f * 1.44269504088896340736f;   // 64^-0.5 * log2(e)
  WF_CHECK_CUDA(launch_pdl(2, latent_attn1_kernel, dim3(B), dim3(L1_THREADS), static_cast<size_t>(smem), stream, mx, mq,
                           reinterpret_cast<__nv_bfloat16*>(ctx), T, H, hp, ns, nslot, sl2));
  count_launch();
  return WF_OK;
}

}  // namespace wf
