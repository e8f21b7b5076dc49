// bf16 linear layer for LARGE M on CTA pairs:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)   (tcgen05.mma cta_group::2)
//
// Same contract as gemm_tc.cu (reference whisper/model.py:35-50 for the encoder / per-clip K,V projections), other
// operand economy.  The single-CTA kernel moves 48 KB of operands into an SM per 128 x 256 x 64 MMA block
// (16 KB of A + 32 KB of W); at the ~64 B/clk an SM can pull from L2 that is ~750 cycles for 512 cycles of tensor
// work - ncu shows the tensor pipe 59 % active.  Here two CTAs of a cluster (the two SMs of a TPC) own one 256 x 256
// tile: each loads ITS 128 rows of A and HALF of the W tile (32 KB per SM per k-block), the leader CTA issues
// 256 x 256 x 16 MMAs that read both halves, and each CTA keeps the accumulator of its own 128 rows in its own TMEM.
//
//   warp 0 lane 0 (both CTAs) : TMA producer; every load signals the LEADER's full barrier
//   warp 1 lane 0 (leader)    : MMA issuer; tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   warp 2       (both CTAs)  : TMEM allocator (cta_group::2)
//   warps 4..11  (both CTAs)  : epilogue of the CTA's own 128 rows (same fused epilogue as gemm_tc.cu); arrives on the
//                               leader's accumulator-free barrier
#include "common.cuh"
#include "kernels.h"
#include "gemm_epilogue.cuh"

namespace wf {

static constexpr int P_BM = 128;       // rows per CTA (the pair covers 256)
static constexpr int P_BN = 256;       // tile width; each CTA stages 128 of the 256 W rows
static constexpr int P_BK = 64;
static constexpr int P_STAGES = 6;
static constexpr int P_A_BYTES = P_BM * P_BK * 2;          // 16 KB
static constexpr int P_B_BYTES = (P_BN / 2) * P_BK * 2;    // 16 KB
static constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
static constexpr int P_ACC_STAGES = 2;
static constexpr int P_TMEM_COLS = P_ACC_STAGES * P_BN;    // 512
static constexpr int P_BAR_BYTES = (2 * P_STAGES + 2 * P_ACC_STAGES) * 8 + 32;
static constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + P_BAR_BYTES + 1024;
static constexpr int P_THREADS = 384;
static constexpr int P_EPI_THREADS = 256;

__device__ __forceinline__ uint32_t p_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t p_cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t p_nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void p_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t p_mapa(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
// TMA tile load whose completion bytes are credited to an mbarrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void p_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                              int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void p_mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void p_umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at the same shared-memory offset in BOTH CTAs
__device__ __forceinline__ void p_umma_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void p_tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(P_TMEM_COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void p_tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(P_TMEM_COLS) : "memory");
}

__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                int M, int N, int K, TcEpilogue ep) {
  extern __shared__ uint8_t p_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + P_STAGES * P_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full_bar = bars;                         // used in the leader only (bytes of both CTAs land here)
  uint64_t* empty_bar = bars + P_STAGES;             // one set per CTA, arrived by the multicast commit
  uint64_t* tfull_bar = bars + 2 * P_STAGES;         // one set per CTA, arrived by the multicast commit
  uint64_t* tempty_bar = bars + 2 * P_STAGES + P_ACC_STAGES;  // leader only: both CTAs' epilogue threads arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 2 * P_ACC_STAGES);

  // Role index.  With the GELU epilogue the eight epilogue warps are busy enough to delay the one-thread producer /
  // issuer roles on their schedulers: a scheduler picks the eligible warp with the HIGHEST hardware id first, so for that
  // epilogue the helper roles move to hardware warps 8-11 (the TMEM lane quarter warp & 3 is the same in both numberings).
  // Measured (M = 192000, K = 1280): N = 5120 with GELU 2242-2252 -> 2186-2214 us; without an activation the plain
  // numbering is 1-2 % faster (1464 vs 1478 us at N = 3840), so it stays for those.
  const int warp = ((threadIdx.x >> 5) + (ep.act != 0 ? 4 : 0)) % (P_THREADS / 32);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = p_cluster_ctarank();
  const bool leader = rank == 0;
  pdl_trigger();

  const int m_tiles = (M + 2 * P_BM - 1) / (2 * P_BM);
  const int n_tiles = (N + P_BN - 1) / P_BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + P_BK - 1) / P_BK;
  const int first = static_cast<int>(p_cluster_id_x());
  const int step = static_cast<int>(p_nclusters_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);    // the leader's producer arrives once and expects the bytes of both CTAs
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < P_ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * P_EPI_THREADS);
    }
    mbar_fence_init();
  }
  if (warp == 2) p_tmem_alloc(tmem_slot);
  tc_fence_before();
  __syncwarp();
  p_cluster_sync();  // barriers of both CTAs initialised, TMEM allocated in both, before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    bool waited = false;
    for (int tile = first; tile < num_tiles; tile += step) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int row_a = m_blk * (2 * P_BM) + static_cast<int>(rank) * P_BM;
      const int row_b = n_blk * P_BN + static_cast<int>(rank) * (P_BN / 2);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t full_leader = p_mapa(smem_u32(&full_bar[stage]), 0);
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
        p_tma_load_2d(smem_b + stage * P_B_BYTES, &map_b, full_leader, kb * P_BK, row_b);
        if (!waited) { pdl_wait(); waited = true; }  // weights above do not depend on the previous kernel; A does
        p_tma_load_2d(smem_a + stage * P_A_BYTES, &map_a, full_leader, kb * P_BK, row_a);
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (!waited) pdl_wait();
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    constexpr uint32_t idesc = umma_idesc_bf16(2 * P_BM, P_BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = first; tile < num_tiles; tile += step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * P_BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * P_A_BYTES));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * P_B_BYTES));
#pragma unroll
        for (int k = 0; k < P_BK / 16; ++k)
          p_umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        p_umma_commit_both(&empty_bar[stage]);
        if (kb == k_blocks - 1) p_umma_commit_both(&tfull_bar[acc]);
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int csel = (warp - 4) >> 2;
    const int rloc = q * 32 + lane;
    pdl_wait();
    long long c_off = 0;
    if (ep.c_off_ptr) c_off = static_cast<long long>(*ep.c_off_ptr) * ep.c_off_mul;
    const float gate = ep.gate ? tanhf(*ep.gate) : 1.0f;
    int it = 0;
    for (int tile = first; tile < num_tiles; tile += step, ++it) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m = m_blk * (2 * P_BM) + static_cast<int>(rank) * P_BM + rloc;
      const bool row_ok = m < M;
      const long long res_row = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
      float mean = 0.f, rstd = 1.f;
      if (ep.ln_colsum && row_ok) {  // issued before the accumulator wait: the loads overlap the main loop
        float s1 = 0.f, s2 = 0.f;
        const float4* sp = reinterpret_cast<const float4*>(ep.stat_in + static_cast<long long>(m) * ep.stat_in_slots * 2);
        const int n4 = ep.stat_in_slots >> 1;
        float4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = (i < n4) ? __ldg(sp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1 += t[i].x + t[i].z;
          s2 += t[i].y + t[i].w;
        }
        for (int i = 8; i < n4; ++i) {
          const float4 u = __ldg(sp + i);
          s1 += u.x + u.z;
          s2 += u.y + u.w;
        }
        mean = s1 / static_cast<float>(K);
        rstd = rsqrtf(fmaxf(s2 / static_cast<float>(K) - mean * mean, 0.f) + ep.ln_eps);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * P_BN;
      float2 st = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = csel; c < P_BN / 32; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(tsrc + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * P_BN + c * 32;
        if (row_ok && n0 < N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          finish_chunk<32>(v, ep, m, res_row, n0, N, gate, c_off, mean, rstd, ep.stat_out ? &st : nullptr);
        }
      }
      if (ep.stat_out && row_ok)
        reinterpret_cast<float2*>(ep.stat_out)[(static_cast<long long>(m) * n_tiles + n_blk) * 2 + csel] = st;
      tc_fence_before();
      p_mbar_arrive_remote(p_mapa(smem_u32(&tempty_bar[acc]), 0));  // this thread is done with the TMEM stage
    }
  }

  // nobody leaves (or frees TMEM) while the peer can still signal its barriers or the pair's MMAs are in flight
  tc_fence_before();
  __syncwarp();
  p_cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    p_tmem_dealloc(tmem_base);
  }
}

bool pair_gemm_usable(int M, int N, int K) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("WF_GEMM_PAIR");  // 0: single-CTA kernel everywhere (A/B measurements)
    mode = e ? atoi(e) : 1;
  }
  return mode != 0 && M >= 2048 && N >= 256 && K >= 64;
}

int linear_bf16_pair(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                     const TcEpilogue& ep, cudaStream_t stream) {
  CUtensorMap ma, mb;
  int rc = make_map_bf16(&ma, A, M, K, lda, P_BM);
  if (rc) return rc;
  rc = make_map_bf16(&mb, W, N, K, ldw, P_BN / 2);
  if (rc) return rc;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
  }
  const int tiles = ((M + 2 * P_BM - 1) / (2 * P_BM)) * ((N + P_BN - 1) / P_BN);
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = 2;
  attr[n].val.clusterDim.y = 1;
  attr[n].val.clusterDim.z = 1;
  ++n;
  if (pdl_enabled(0)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  WF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel, ma, mb, M, N, K, ep));
  count_launch();
  return WF_OK;
}

}  // namespace wf
