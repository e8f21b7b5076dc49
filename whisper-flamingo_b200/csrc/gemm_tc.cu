// bf16 linear layer on the 5th-gen tensor cores:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// Replaces reference whisper/model.py:35-41 (Linear.forward -> F.linear -> cuBLAS) and the
// conv stem (model.py:44-50, after im2col) for the bf16 engine.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor, 128B swizzle, OOB rows/cols zero-filled,
//                                    so M, N and K tails need no padding)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16 per instruction,
//                                    fp32 accumulators in TMEM, double-buffered across tiles)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue       (tcgen05.ld 32x32b -> bias / exact-erf GELU / tanh(gate) / residual
//                                    -> bf16 or fp32 stores)
// Tile order is n-fastest so the CTAs of a wave share the same A rows through L2 while the whole
// weight matrix (<= 13 MB) stays L2-resident: A is streamed from HBM once per GEMM.
#include "common.cuh"
#include "kernels.h"
#include "gemm_epilogue.cuh"

namespace wf {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom
static constexpr int UMMA_K = 16;

template <int BN, int STAGES>
struct TcCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BN;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = (2 * STAGES + 2 * ACC_STAGES) * 8 + 32;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static_assert(TMEM_COLS_RAW <= 512, "accumulators do not fit TMEM");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-B alignment");
};

static constexpr int TC_THREADS = 384;      // 4 control warps + 8 epilogue warps
static constexpr int EPI_THREADS = 256;
static constexpr int SPLITK_COUNTER_BYTES = 4096;  // 1024 per-tile arrival counters at the head of the workspace

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               int M, int N, int K, TcEpilogue ep) {
  using Cfg = TcCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + Cfg::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * Cfg::ACC_STAGES);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();  // the next kernel in the stream may start its own prologue now

  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + BK - 1) / BK;
  const int S = ep.splits;
  const int kb_per = (k_blocks + S - 1) / S;
  const int num_items = num_tiles * S;  // item = tile * S + split (splits of a tile run on neighbouring CTAs)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_THREADS);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    // The weight (B) tiles of the first stages do not depend on the previous kernel: they are requested before the
    // programmatic-dependency wait, so their DRAM latency overlaps the predecessor's tail; A tiles follow the wait.
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / S, split = item - tile * S;
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      int kb = kb0;
      if (first) {
        first = false;
        const int npre = min(STAGES, kb1 - kb0);
        for (int i = 0; i < npre; ++i) {
          mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
          tma_load_2d(smem_b + i * Cfg::B_BYTES, &map_b, &full_bar[i], (kb0 + i) * BK, n_blk * BN);
        }
        pdl_wait();
        for (int i = 0; i < npre; ++i)
          tma_load_2d(smem_a + i * Cfg::A_BYTES, &map_a, &full_bar[i], (kb0 + i) * BK, m_blk * BM);
        kb = kb0 + npre;
        if (npre == STAGES) { stage = 0; phase = 1; } else { stage = npre; }
      }
      for (; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_2d(smem_a + stage * Cfg::A_BYTES, &map_a, &full_bar[stage], kb * BK, m_blk * BM);
        tma_load_2d(smem_b + stage * Cfg::B_BYTES, &map_b, &full_bar[stage], kb * BK, n_blk * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int split = item % S;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps; warp & 3 = TMEM lane quarter,
    // the two warps of a quarter take alternate 32-column chunks
    const int q = warp & 3;
    const int csel = (warp - 4) >> 2;
    const int rloc = q * 32 + lane;
    pdl_wait();  // residual / gate / offset / workspace may be produced by the previous kernel
    long long c_off = 0;
    if (ep.c_off_ptr) c_off = static_cast<long long>(*ep.c_off_ptr) * ep.c_off_mul;
    const float gate = ep.gate ? tanhf(*ep.gate) : 1.0f;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int tile = item / S;
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + rloc;
      const bool row_ok = m < M;
      const long long res_row = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
      const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (S == 1) {
        float mean = 0.f, rstd = 1.f;
        if (ep.ln_colsum && row_ok) {  // LayerNorm statistics of this row, emitted by the kernel that produced it
          // all partials of the row are requested before the first add (a counted loop would serialise the round trips)
          float s1 = 0.f, s2 = 0.f;
          const float4* sp = reinterpret_cast<const float4*>(ep.stat_in + static_cast<long long>(m) * ep.stat_in_slots * 2);
          const int n4 = ep.stat_in_slots >> 1;  // slots are written in pairs (two column halves per n-tile)
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
        float2 st = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = csel; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(tsrc + c * 32, r);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
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
        mbar_arrive(&tempty_bar[acc]);
      } else {
        // ---- split-K: publish this slice's fp32 partial, the last slice to arrive reduces and finishes the tile
        float* part = ep.ws_part + static_cast<long long>(item) * (BM * BN) + rloc * BN;
#pragma unroll 1
        for (int c = csel; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(tsrc + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(part + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);  // TMEM stage is free again
        __threadfence();
        epi_bar();
        if (threadIdx.x == 4 * 32) {
          const int prev = atomicAdd(ep.ws_count + tile, 1);
          *last_flag = (prev == S - 1);
        }
        epi_bar();
        if (*last_flag) {
          __threadfence();
          const float* base = ep.ws_part + static_cast<long long>(tile) * S * (BM * BN) + rloc * BN;
#pragma unroll 1
          for (int c = csel; c < BN / 32; c += 2) {
            const int n0 = n_blk * BN + c * 32;
            if (row_ok && n0 < N) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
              for (int s2 = 0; s2 < S; ++s2) {
                const float* p = base + static_cast<long long>(s2) * (BM * BN) + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 t4 = __ldcg(reinterpret_cast<const float4*>(p + j));
                  v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
                }
              }
              finish_chunk<32>(v, ep, m, res_row, n0, N, gate, c_off);
            }
          }
          if (threadIdx.x == 4 * 32) ep.ws_count[tile] = 0;  // self-resetting: ready for the next launch
        }
        epi_bar();  // last_flag is rewritten only after everyone has read it
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D row-major bf16 matrix [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return WF_ERR_CUDA;
  }
  WF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  WF_REQUIRE((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes (ld=%lld)", ld);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
    return WF_ERR_CUDA;
  }
  return WF_OK;
}

template <int BN, int STAGES>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, TcEpilogue ep,
                     void* ws, long long ws_bytes, cudaStream_t stream) {
  using Cfg = TcCfg<BN, STAGES>;
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int k_blocks = (K + BK - 1) / BK;
  // split-K only pays for long-K weight-streaming shapes that cannot fill the machine with output tiles
  // (measured on B200, M=128: K=5120 28 -> 22 us; at K=1280 the partial-tile round trip costs more than it saves)
  int S = 1;
  if (ws && M <= 2 * BM && tiles <= 1024 && tiles * 4 <= 3 * num_sms() && k_blocks >= 64) {
    S = (num_sms() + tiles - 1) / tiles;
    if (S > k_blocks / 4) S = k_blocks / 4;
    if (S > 8) S = 8;
    while (S > 1 && SPLITK_COUNTER_BYTES + static_cast<long long>(tiles) * S * BM * BN * 4 > ws_bytes) --S;
    if (S > 1) {  // every slice must own at least one k-block
      const int per = (k_blocks + S - 1) / S;
      S = (k_blocks + per - 1) / per;
    }
    if (S < 1) S = 1;
  }
  ep.splits = S;
  ep.ws_count = reinterpret_cast<int*>(ws);
  ep.ws_part = ws ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + SPLITK_COUNTER_BYTES) : nullptr;
  const int items = tiles * S;
  const int grid = items < num_sms() ? items : num_sms();
  WF_CHECK_CUDA(launch_pdl(0, gemm_tc_kernel<BN, STAGES>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, ma, mb, M,
                           N, K, ep));
  count_launch();
  return WF_OK;
}

int linear_bf16_tc(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                   const LinearEpilogue& e, int tile_hint, cudaStream_t stream) {
  WF_REQUIRE(M > 0 && N > 0 && K > 0, "linear: empty problem M=%d N=%d K=%d", M, N, K);
  int bn = tile_hint;
  if (bn == 0) {
    if (M <= 2048) {
      // few row tiles (beam-search decode steps, short prefills): the widest tile that still gives every SM a tile
      const int m_tiles = (M + BM - 1) / BM;
      bn = 32;
      for (int cand = (M <= 256 ? 128 : 256); cand > 32; cand >>= 1)  // M <= 256 streams weights: measured 128 > 256
        if (static_cast<long long>(m_tiles) * ((N + cand - 1) / cand) >= num_sms()) { bn = cand; break; }
    } else {
      bn = N >= 256 ? 256 : 128;
    }
  }
  WF_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "linear: unsupported tile hint %d", tile_hint);
  WF_REQUIRE(e.split_n >= 0 && e.split_n % 64 == 0 && e.split_n < N && (e.split_n == 0 || (e.C2 && e.hm_heads > 0)),
             "linear: bad two-destination spec (split_n=%d N=%d)", e.split_n, N);
  if (e.hm_heads > 0) {
    const int n_hm = N - e.split_n;
    WF_REQUIRE(n_hm % 64 == 0 && e.hm_heads * 64 == n_hm && e.hm_T > 0 && e.hm_rpb > 0 && !e.residual,
               "linear: bad head-major output spec (N=%d heads=%d T=%d rpb=%d)", n_hm, e.hm_heads, e.hm_T, e.hm_rpb);
  }
  WF_REQUIRE(!e.ws || ((reinterpret_cast<uintptr_t>(e.ws) & 15) == 0 && e.ws_bytes >= SPLITK_COUNTER_BYTES),
             "linear: split-K workspace must be 16-byte aligned and hold at least %d bytes", SPLITK_COUNTER_BYTES);
  TcEpilogue ep;
  ep.C = e.C; ep.ldc = e.ldc; ep.bias = e.bias; ep.residual = e.residual; ep.ldr = e.ldr;
  ep.res_row_mod = e.res_row_mod; ep.gate = e.gate; ep.act = e.act; ep.out_f32 = e.out_f32;
  ep.c_off_ptr = e.c_off_ptr; ep.c_off_mul = e.c_off_mul;
  ep.hm_heads = e.hm_heads; ep.hm_T = e.hm_T; ep.hm_rpb = e.hm_rpb;
  ep.splits = 1; ep.ws_part = nullptr; ep.ws_count = nullptr;
  ep.ln_colsum = e.ln_colsum; ep.ln_eps = e.ln_eps; ep.split_n = e.split_n; ep.C2 = e.C2;
  ep.stat_out = e.stat_out; ep.stat_in = e.stat_in; ep.stat_in_slots = e.stat_in_slots;
  {
    int sbn = 0, scs = 1, sbm = 128;  // decode step (a few row tiles): cluster split-K kernel
    // (row statistics travelling between GEMMs are a feature of the large-tile kernels: those calls stay there)
    const bool stats = e.stat_out != nullptr || e.stat_in != nullptr;
    if (!(stats && M > 128) && skinny_plan(M, N, K, tile_hint, &sbn, &scs, &sbm))
      return linear_bf16_skinny(A, lda, W, ldw, M, N, K, ep, sbm, sbn, scs, stream);
    // a LayerNorm computed inside the GEMM exists in the skinny kernel only (the large-tile kernels take the row
    // statistics from the producing GEMM): more than one wave of 128-wide tiles rather than no plan
    if (e.ln_colsum && !e.stat_in && M <= 512) return linear_bf16_skinny(A, lda, W, ldw, M, N, K, ep, 128, 128, 1, stream);
  }
  WF_REQUIRE(!e.ln_colsum || (e.stat_in && e.stat_in_slots > 0),
             "linear: a fused LayerNorm over more than 128 rows needs the row statistics of the producing GEMM (M=%d)", M);
  WF_REQUIRE(!(e.stat_out && e.ws), "linear: row statistics cannot be combined with the split-K workspace");
  if ((tile_hint == 0 || tile_hint == 256) && !e.ws && pair_gemm_usable(M, N, K))
    return linear_bf16_pair(A, lda, W, ldw, M, N, K, ep, stream);  // large M: 256 x 256 tiles on CTA pairs
  CUtensorMap ma, mb;
  int rc = make_map_bf16(&ma, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_map_bf16(&mb, W, N, K, ldw, bn);
  if (rc) return rc;
  switch (bn) {
    case 32: return launch_tc<32, 8>(ma, mb, M, N, K, ep, e.ws, e.ws_bytes, stream);
    case 64: return launch_tc<64, 8>(ma, mb, M, N, K, ep, e.ws, e.ws_bytes, stream);
    case 128: return launch_tc<128, 6>(ma, mb, M, N, K, ep, e.ws, e.ws_bytes, stream);
    default: return launch_tc<256, 4>(ma, mb, M, N, K, ep, e.ws, e.ws_bytes, stream);
  }
}

}  // namespace wf
