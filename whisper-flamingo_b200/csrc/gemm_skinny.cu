// Decode-step linear layers (M <= 128 rows = one batch of hypotheses):  C[M,N] = epilogue(LN?(A)[M,K] . W[N,K]^T)
//
// Replaces reference whisper/model.py:35-41 (Linear.forward) and, when asked, the LayerNorm in front of it
// (model.py:30-32) for the one-token decoder pass (decoding.py:155-164).  These GEMMs move 3-13 MB of weights and
// should take a microsecond; the persistent large-tile kernel (gemm_tc.cu) needed 8-26 us for them because ONE
// CTA walked the whole K extent of its N tile serially.  Measured on B200 (tools/microbench.py skinny2 / skinny3):
//     t = 1.9 us + 0.30 us per 64-wide k-block, whatever M (16 / 64 / 128 rows, or 64-row MMAs), the N tile (32 .. 128)
//     or where the weights come from (L2 / HBM);
// the cost sits in the single MMA-ISSUING THREAD: ~350 cycles per barrier round (wait, fence, commit) + ~63 cycles per
// tcgen05.mma, while the TMA stream alone needs 0.155 us per 20 KB k-block (~64 B/clk per SM).
//
// So this kernel (a) issues MMAs from TWO warps into two TMEM accumulators, (b) moves two swizzle atoms per barrier
// round and (c) can split K as well as N: a thread-block CLUSTER of CS CTAs owns one 128 x BN output tile, each CTA
// accumulates a K slice in TMEM with two swizzle atoms (K = 128) per pipeline round, and the partial tiles are reduced
// through distributed shared memory: every CTA dumps its accumulator to its own smem, cluster barrier, then CTA r
// sums row slice r of all CS partials (coalesced ld.shared::cluster) and runs the fused epilogue for it.  No global
// workspace, no atomics, no second kernel.
//
// Optional fused LayerNorm: the epilogue warps, idle during the main loop, accumulate sum / sum-of-squares of the raw
// A rows straight from the TMA-staged smem tiles (a 128-byte swizzle permutes 16-byte chunks inside a row, which a
// row sum does not care about); the row statistics travel with the partial tiles through DSMEM and the epilogue
// applies  y = rstd * (acc - mean * colsum(W')) + bias'.
#include "common.cuh"
#include "kernels.h"
#include "gemm_epilogue.cuh"

namespace wf {

static constexpr int SK_BM = 128;
static constexpr int SK_BK = 64;       // one 128-byte swizzle atom of bf16
static constexpr int SK_KA = 2;        // atoms per pipeline round
static constexpr int SK_THREADS = 384; // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4..11 epilogue / LN stats
static constexpr int SK_UMMA_K = 16;
static constexpr int EPI_THREADS_SK = 256;

// BM = 128, or 64: half the rows per CTA on twice the CTAs.  Every CTA of a GEMM re-reads its BM rows of the activation
// tile from L2 (~64 B/clk per SM, ~6300 B/clk chip-wide), which - not the weight stream - sets the main loop of these
// kernels: (BM + BN) x K x 2 bytes per CTA.  N = K = 1280: 409 KB on 40 CTAs (128 x 32) -> 245 KB on 80 CTAs (64 x 32).
template <int BM, int BN, int STAGES>
struct SkCfg {
  static constexpr int A_ATOM = BM * SK_BK * 2;  // 16 KB / 8 KB
  static constexpr int B_ATOM = BN * SK_BK * 2;
  static constexpr int STAGE_BYTES = SK_KA * (A_ATOM + B_ATOM);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int DUMP_LD = BN + 4;  // +16 B per row: 16-byte accesses of a quarter-warp hit distinct banks
  static constexpr int DUMP_BYTES = BM * DUMP_LD * 4;
  static constexpr int STAT_BYTES = 2 * SK_BM * 2 * 4;  // [atom group][row]{sum, sumsq}
  static constexpr int BAR_BYTES = (2 * STAGES + 1) * 8 + 16;
  static constexpr int SMEM_BYTES = PIPE_BYTES + STAT_BYTES + BAR_BYTES + 1024;
  // MMA-issuing warps, one TMEM accumulator each.  Measured (N=K=1280, 32-wide tile): 1 -> 6.9 us, 2 -> 5.9 us, 3 -> 6.1 us:
  // with two streams the main loop runs at the ~64 B/clk an SM pulls through the crossbar (0.33 us per 40 KB round).
  static constexpr int NI = 2;
  static constexpr int TMEM_RAW = NI * BN;
  static constexpr int TMEM_COLS = TMEM_RAW <= 32 ? 32 : TMEM_RAW <= 64 ? 64 : TMEM_RAW <= 128 ? 128 : TMEM_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS <= 512, "accumulators do not fit tensor memory");
  static_assert(PIPE_BYTES >= DUMP_BYTES, "the accumulator dump aliases the pipeline stages");
  static_assert(B_ATOM % 1024 == 0, "operand atoms must keep 1024-byte alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "over the shared-memory budget");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}

template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   int M, int N, int K, TcEpilogue ep, int CS) {
  using Cfg = SkCfg<BM, BN, STAGES>;
  extern __shared__ uint8_t sk_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk_smem_raw) + 1023) & ~uintptr_t(1023));
  float* dump = reinterpret_cast<float*>(smem);                                   // aliases the pipeline stages
  float* stat = reinterpret_cast<float*>(smem + Cfg::PIPE_BYTES);                 // [2][128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + Cfg::STAT_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  const int rank = static_cast<int>(cluster_ctarank());
  const int n_blk = blockIdx.x / CS;
  const int m0 = blockIdx.y * BM;   // row tile: beam-search steps run up to a few hundred hypotheses
  const int rounds_total = (K + SK_KA * SK_BK - 1) / (SK_KA * SK_BK);
  const int per = (rounds_total + CS - 1) / CS;
  const int r0 = min(rounds_total, rank * per), r1 = min(rounds_total, r0 + per);
  const int nr = r1 - r0;  // pipeline rounds of this CTA (0 is legal: it contributes a zero partial)
  const bool ln = ep.ln_colsum != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], ln ? 9 : 1);  // MMA commit (+ one arrival per LN-statistics warp)
    }
    mbar_init(tfull_bar, nr < Cfg::NI ? (nr > 0 ? nr : 1) : Cfg::NI);  // one commit per MMA-issuing warp that has rounds
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto a_atom = [&](int stage, int a) { return smem + stage * Cfg::STAGE_BYTES + a * Cfg::A_ATOM; };
  auto b_atom = [&](int stage, int a) { return smem + stage * Cfg::STAGE_BYTES + SK_KA * Cfg::A_ATOM + a * Cfg::B_ATOM; };

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer: weights first (they do not depend on
    // the previous kernel), activations after the programmatic-dependency wait
    const int npre = min(STAGES, nr);
    for (int i = 0; i < npre; ++i) {
      mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
#pragma unroll
      for (int a = 0; a < SK_KA; ++a)
        tma_load_2d(b_atom(i, a), &map_b, &full_bar[i], ((r0 + i) * SK_KA + a) * SK_BK, n_blk * BN);
    }
    pdl_wait();
    for (int i = 0; i < npre; ++i) {
#pragma unroll
      for (int a = 0; a < SK_KA; ++a)
        tma_load_2d(a_atom(i, a), &map_a, &full_bar[i], ((r0 + i) * SK_KA + a) * SK_BK, m0);
    }
    int stage = npre == STAGES ? 0 : npre;
    uint32_t phase = npre == STAGES ? 1 : 0;
    for (int r = npre; r < nr; ++r) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
      for (int a = 0; a < SK_KA; ++a) {
        tma_load_2d(a_atom(stage, a), &map_a, &full_bar[stage], ((r0 + r) * SK_KA + a) * SK_BK, m0);
        tma_load_2d(b_atom(stage, a), &map_b, &full_bar[stage], ((r0 + r) * SK_KA + a) * SK_BK, n_blk * BN);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 1 && warp <= Cfg::NI && lane == 0) {
    // ------------------------------------------------------------ MMA issuers: warp 1 takes the even pipeline rounds
    // into accumulator 0, warp 3 the odd rounds into accumulator 1 (the epilogue adds them).  One thread sustains
    // ~350 cycles per round + ~63 per MMA whatever the tile shape (M = 64 operands: same time), and the cost is per issuing
    // THREAD: a second stream brought N=K=1280 from 6.9 to 5.9 us and K=5120 from 20.0 to 15.4 us.
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    constexpr int NI = Cfg::NI;
    const int me = warp - 1;
    const uint32_t tmem_d = tmem_base + me * BN;
    int stage = me % STAGES;
    uint32_t phase = 0;
    int last = -1;
    for (int r = me; r < nr; r += NI) last = r;
    for (int r = me; r < nr; r += NI) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
#pragma unroll
      for (int a = 0; a < SK_KA; ++a) {
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(a_atom(stage, a)));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(b_atom(stage, a)));
#pragma unroll
        for (int k = 0; k < SK_BK / SK_UMMA_K; ++k)
          umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (r > me || a > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);
      if (r == last) umma_commit(tfull_bar);
      stage += NI;
      while (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps, phase A
    const int q = warp & 3;            // TMEM lane quarter
    const int grp = (warp - 4) >> 2;   // 0 / 1: column half in the epilogue, swizzle atom in the LN statistics
    const int rloc = q * 32 + lane;    // row of the staged A tile whose statistics this thread accumulates
    // accumulator row of this thread: a 128-row accumulator keeps row r in lane r, a 64-row one keeps rows 16 q .. 16 q + 15
    // in lanes 32 q .. 32 q + 15 (profiles/r01_probe_tmem_m64_layout.txt)
    const int arow = BM == 128 ? rloc : q * 16 + (lane & 15);
    const bool alive = BM == 128 || lane < 16;
    if (ln) {
      // row sums of the raw A operand from the staged tiles (row = 128 B = 8 chunks of 16 B, order irrelevant)
      float s1 = 0.f, s2 = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int r = 0; r < nr; ++r) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* row = a_atom(stage, grp) + rloc * 128;
#pragma unroll
        for (int j = 0; j < 8 && rloc < BM; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(row + ((j + rloc) & 7) * 16);
          const float f[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y),
                              bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            s1 += f[e];
            s2 = fmaf(f[e], f[e], s2);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      *reinterpret_cast<float2*>(stat + (grp * SK_BM + rloc) * 2) = make_float2(s1, s2);
    }
    if (nr > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    if (CS > 1) {
      // the dump aliases the pipeline stages: every LayerNorm-statistics warp must have finished reading the A tiles
      if (ln) epi_bar();
      // all TMA loads of this CTA have been consumed and all its MMAs have retired: the stages are free to hold the dump
      const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = grp; c < BN / 32; c += 2) {
        uint32_t r[32];
        if (nr > 0) {
          tmem_ld_32x32(tsrc + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int a2 = 1; a2 < Cfg::NI; ++a2) {  // rounds were dealt to NI accumulators
            if (a2 < nr) {
              uint32_t r2[32];
              tmem_ld_32x32(tsrc + a2 * BN + c * 32, r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        float* drow = dump + arow * Cfg::DUMP_LD + c * 32;
        if (alive) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(drow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
      }
    }
  }

  // ---------------------------------------------------------------- partial tiles (and row statistics) are published
  __syncwarp();
  if (CS > 1) cluster_sync_all();
  else if (ln) __syncthreads();

  if (warp >= 4) {
    pdl_wait();  // residual / gate / offset may be produced by the previous kernel
    long long c_off = 0;
    if (ep.c_off_ptr) c_off = static_cast<long long>(*ep.c_off_ptr) * ep.c_off_mul;
    const float gate = ep.gate ? tanhf(*ep.gate) : 1.0f;
    const uint32_t st_local = smem_u32(stat);
    // per-row LayerNorm statistics: partial sums of every CTA of the cluster and of both atom groups
    auto row_stats = [&](int row, float& mean, float& rstd) {
      float2 t[8][2];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (s < CS) {
#pragma unroll
          for (int g = 0; g < 2; ++g)
            t[s][g] = ld_dsmem_f2(dsmem_addr(st_local + static_cast<uint32_t>((g * SK_BM + row) * 8), s));
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (s < CS) {
          s1 += t[s][0].x + t[s][1].x;
          s2 += t[s][0].y + t[s][1].y;
        }
      }
      mean = s1 / static_cast<float>(K);
      rstd = rsqrtf(fmaxf(s2 / static_cast<float>(K) - mean * mean, 0.f) + ep.ln_eps);
    };
    if (CS == 1) {
      const int q = warp & 3;
      const int grp = (warp - 4) >> 2;
      const int ml = BM == 128 ? q * 32 + lane : q * 16 + (lane & 15);       // row inside the tile
      const bool alive = BM == 128 || lane < 16;
      const int m = m0 + ml;
      const long long res_row = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
      float mean = 0.f, rstd = 1.f;
      if (ln) row_stats(ml, mean, rstd);
      const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = grp; c < BN / 32; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(tsrc + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int a2 = 1; a2 < Cfg::NI; ++a2) {  // rounds were dealt to NI accumulators
          if (a2 < nr) {
            uint32_t r2[32];
            tmem_ld_32x32(tsrc + a2 * BN + c * 32, r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
        }
        const int n0 = n_blk * BN + c * 32;
        if (alive && m < M && n0 < N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          finish_chunk<32>(v, ep, m, res_row, n0, N, gate, c_off, mean, rstd);
        }
      }
    } else {
      // CTA `rank` finishes rows [rank * 128 / CS, +128 / CS) of the tile.  A thread takes 8 consecutive columns, so a
      // warp reads 1 KB of CONTIGUOUS remote shared memory per partial tile (a row-per-thread mapping measured 8 GB/s
      // over DSMEM: 32 different 128-byte lines per request) and stores coalesced 16-byte pieces of output rows.
      const int et = (warp - 4) * 32 + lane;
      const int rows_per = (BM + CS - 1) / CS;   // CS need not divide the tile height (clusters of 5 or 6)
      constexpr int GPR = BN / 8;
      const uint32_t dump_local = smem_u32(dump);
      for (int it = et; it < rows_per * GPR; it += EPI_THREADS_SK) {
        const int row = rank * rows_per + it / GPR;     // row inside the tile
        const int col = (it % GPR) * 8;
        const int n0 = n_blk * BN + col;
        if (row >= BM || m0 + row >= M || n0 >= N) continue;
        const uint32_t off = static_cast<uint32_t>((row * Cfg::DUMP_LD + col) * 4);
        float4 t[8][2];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < CS) {
            const uint32_t base = dsmem_addr(dump_local + off, static_cast<uint32_t>(s));
            t[s][0] = ld_dsmem_f4(base);
            t[s][1] = ld_dsmem_f4(base + 16);
          }
        }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < CS) {
            v[0] += t[s][0].x; v[1] += t[s][0].y; v[2] += t[s][0].z; v[3] += t[s][0].w;
            v[4] += t[s][1].x; v[5] += t[s][1].y; v[6] += t[s][1].z; v[7] += t[s][1].w;
          }
        }
        float mean = 0.f, rstd = 1.f;
        if (ln) row_stats(row, mean, rstd);
        const long long res_row = ep.res_row_mod > 0 ? ((m0 + row) % ep.res_row_mod) : (m0 + row);
        finish_chunk<8>(v, ep, m0 + row, res_row, n0, N, gate, c_off, mean, rstd);
      }
    }
  }

  // nobody may leave (and release its shared memory) while a peer can still read it
  __syncwarp();
  if (CS > 1) cluster_sync_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
template <int BM, int BN, int STAGES>
static int launch_skinny(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const TcEpilogue& ep, int cs,
                         cudaStream_t stream) {
  using Cfg = SkCfg<BM, BN, STAGES>;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<BM, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
  }
  const int tiles = (N + BN - 1) / BN;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * cs, (M + BM - 1) / BM);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = cs;
  attr[n].val.clusterDim.y = 1;
  attr[n].val.clusterDim.z = 1;
  ++n;
  if (pdl_enabled(0)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  const int prio = launch_priority(0);
  if (prio != INT_MIN) {
    attr[n].id = cudaLaunchAttributePriority;
    attr[n].val.priority = prio;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  WF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<BM, BN, STAGES>, ma, mb, M, N, K, ep, cs));
  count_launch();
  return WF_OK;
}

// Tile width and cluster size for an [M <= 128] x N x K problem.  Measured on B200 inside a CUDA graph with PDL
// (tools/microbench.py skinny3, M = 128, two MMA-issuing warps):
//   * fixed cost ~2.3 us; one CTA then walks a 128-wide K round (40 KB for a 32-wide tile) in ~0.33 us, the ~64 B/clk an
//     SM pulls through the crossbar;
//   * a cluster costs ~1.1 us for its two barriers plus the DSMEM reduction at ~20 B/clk per SM, so K is split (4-way)
//     only when a CTA would otherwise walk >= 32 rounds:  N=1280, K=5120: 15.4 us (32, cs 1) -> 10.0 us (64, cs 4);
//     N=K=1280: 5.94 us (32, cs 1) vs 5.98 us (32, cs 2).
// M > 128 (beam search: B x G hypotheses): the same kernel over ceil(M / 128) row tiles, every tile a CTA (cluster) of
// its own - the weights are read once from HBM and then from L2.  Up to 4 row tiles; above, the persistent large-tile
// kernels of gemm_tc.cu / gemm_tc2.cu take over.
static constexpr int SK_MAX_M_TILES = 4;
bool skinny_plan(int M, int N, int K, int tile_hint, int* bn_out, int* cs_out, int* bm_out) {
  if (M > SK_BM * SK_MAX_M_TILES) return false;
  const int m_tiles = (M + SK_BM - 1) / SK_BM;
  static int mode = -1, force_cs = -1, force_bm = -1;
  if (mode < 0) {
    const char* e = getenv("WF_SKINNY");  // 0: use the persistent kernel of gemm_tc.cu instead (A/B measurements)
    mode = e ? atoi(e) : 1;
    const char* c = getenv("WF_SKINNY_CS");
    force_cs = c ? atoi(c) : 0;
    const char* b = getenv("WF_SKINNY_BM");  // 64 / 128: force the row tile (A/B measurements)
    force_bm = b ? atoi(b) : 0;
  }
  if (mode == 0) return false;
  const int sms = num_sms();
  const int rounds = (K + SK_KA * SK_BK - 1) / (SK_KA * SK_BK);
  const int bns[4] = {32, 64, 128, 256};
  // cluster sizes to try, widest first.  6 CTAs need not divide the tile height (the row slices of the reduction are
  // ceil(128 / 6) rows) and 22 such clusters fit the chip (profiles/r02_probe_cluster_occupancy.txt): N = 1280, K = 5120
  // 9.87 us (cs 4, 80 CTAs) -> 9.57 us (cs 6, 120 CTAs); 5: 9.62, 3: 10.44, 8 does not fit (160 CTAs)
  int cand[5] = {6, 4, 2, 1, 0};
  if (rounds < 32) { cand[0] = 1; cand[1] = 0; }
  if (force_cs) { cand[0] = force_cs; cand[1] = force_cs > 1 ? 1 : 0; cand[2] = 0; }
  bool found = false;
  for (int ci = 0; cand[ci] != 0 && !found; ++ci) {
    const int cs = cand[ci];
    if (cs > 1 && cs > rounds) continue;
    for (int i = 0; i < 4 && !found; ++i) {
      const int bn = bns[i];
      if (tile_hint && bn != tile_hint) continue;
      if (cs == 6 && bn > 64 && !tile_hint) continue;   // measured only where it keeps the narrow tiles
      const int tiles = (N + bn - 1) / bn;
      if (tiles * cs * m_tiles > sms) continue;
      // clusters are placed inside a GPC: on the 148-SM part 33 clusters of 4 and 22 of 6 are resident at once
      // (tools/probe/cluster_occ.cu); more would run as a second wave
      const int max_clusters = cs == 6 ? sms * 22 / 148 : cs == 4 ? sms * 33 / 148 : sms / cs;
      if (tiles * m_tiles > max_clusters) continue;
      *bn_out = bn;
      *cs_out = cs;
      *bm_out = SK_BM;
      found = true;
    }
  }
  // 64-row tiles on twice the CTAs (no K split): taken when a CTA then pulls fewer bytes, (BM + BN) x K x 2
  if (force_bm != 128 && (!found || *cs_out == 1)) {
    const int m_tiles64 = (M + 63) / 64;
    for (int i = 0; i < 3; ++i) {
      const int bn = bns[i];
      if (tile_hint && bn != tile_hint) continue;
      const int tiles = (N + bn - 1) / bn;
      if (tiles * m_tiles64 > sms) continue;
      if (!found || M <= 64 || 64 + bn < SK_BM + *bn_out || force_bm == 64) {
        *bn_out = bn;
        *cs_out = 1;
        *bm_out = 64;
        found = true;
      }
      break;
    }
  }
  return found;
}

int linear_bf16_skinny(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                       const TcEpilogue& ep, int bm, int bn, int cs, cudaStream_t stream) {
  CUtensorMap ma, mb;
  int rc = make_map_bf16(&ma, A, M, K, lda, bm);
  if (rc) return rc;
  rc = make_map_bf16(&mb, W, N, K, ldw, bn);
  if (rc) return rc;
  if (bm == 64) {
    WF_REQUIRE(cs == 1, "linear (skinny): 64-row tiles do not split K");
    switch (bn) {
      case 32: return launch_skinny<64, 32, 4>(ma, mb, M, N, K, ep, cs, stream);
      case 64: return launch_skinny<64, 64, 4>(ma, mb, M, N, K, ep, cs, stream);
      case 128: return launch_skinny<64, 128, 3>(ma, mb, M, N, K, ep, cs, stream);
      default:
        set_error("linear (skinny): unsupported tile 64 x %d", bn);
        return WF_ERR_UNSUPPORTED;
    }
  }
  static int small = -1;  // WF_SKINNY_SMALL=1: <= 100 KB of shared memory per CTA, so that a GEMM CTA fits on an SM next to
  if (small < 0) {        // one K/V-streaming attention CTA of another sub-batch (SplitSession)
    const char* e = getenv("WF_SKINNY_SMALL");
    small = e ? atoi(e) : 0;
  }
  if (small && bn == 32) return launch_skinny<128, 32, 2>(ma, mb, M, N, K, ep, cs, stream);
  if (small && bn == 64) return launch_skinny<128, 64, 2>(ma, mb, M, N, K, ep, cs, stream);
  switch (bn) {
    case 32: return launch_skinny<128, 32, 4>(ma, mb, M, N, K, ep, cs, stream);
    case 64: return launch_skinny<128, 64, 4>(ma, mb, M, N, K, ep, cs, stream);
    case 128: return launch_skinny<128, 128, 3>(ma, mb, M, N, K, ep, cs, stream);
    case 256: return launch_skinny<128, 256, 2>(ma, mb, M, N, K, ep, cs, stream);
    default:
      set_error("linear (skinny): unsupported tile %d", bn);
      return WF_ERR_UNSUPPORTED;
  }
}

}  // namespace wf
