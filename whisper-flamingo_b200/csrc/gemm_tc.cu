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
  static constexpr int BAR_BYTES = (2 * STAGES + 2 * ACC_STAGES) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static_assert(TMEM_COLS_RAW <= 512, "accumulators do not fit TMEM");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-B alignment");
};

struct TcEpilogue {
  void* C;
  long long ldc;
  const float* bias;
  const void* residual;
  long long ldr;
  int res_row_mod;
  const float* gate;
  int act;
  int out_f32;
  const int* c_off_ptr;
  long long c_off_mul;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(256, 1)
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + BK - 1) / BK;

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
      mbar_init(&tempty_bar[i], 128);
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
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      for (int kb = 0; kb < k_blocks; ++kb) {
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads = 128 TMEM lanes)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    long long c_off = 0;
    if (ep.c_off_ptr) c_off = static_cast<long long>(*ep.c_off_ptr) * ep.c_off_mul;
    const float gate = ep.gate ? tanhf(*ep.gate) : 1.0f;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + q * 32 + lane;
      const bool row_ok = m < M;
      const long long res_row = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * BN + c * 32;
        if (row_ok && n0 < N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const bool full = (n0 + 32 <= N);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (full || n0 + j < N) ? __ldg(ep.bias + n0 + j) : 0.f;
          }
          if (ep.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if (ep.gate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= gate;
          }
          if (ep.out_f32) {
            float* crow = reinterpret_cast<float*>(ep.C) + c_off + static_cast<long long>(m) * ep.ldc + n0;
            if (ep.residual) {
              const float* rrow = reinterpret_cast<const float*>(ep.residual) + res_row * ep.ldr + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (full || n0 + j < N) v[j] += rrow[j];
            }
            if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + j < N) crow[j] = v[j];
            }
          } else {
            __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.C) + c_off + static_cast<long long>(m) * ep.ldc + n0;
            if (ep.residual) {
              const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + res_row * ep.ldr + n0;
              if (full && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 u = *reinterpret_cast<const uint4*>(rrow + j);
                  v[j + 0] += bf16lo(u.x); v[j + 1] += bf16hi(u.x);
                  v[j + 2] += bf16lo(u.y); v[j + 3] += bf16hi(u.y);
                  v[j + 4] += bf16lo(u.z); v[j + 5] += bf16hi(u.z);
                  v[j + 6] += bf16lo(u.w); v[j + 7] += bf16hi(u.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (n0 + j < N) v[j] += __bfloat162float(rrow[j]);
              }
            }
            if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16(v[j + 0], v[j + 1]);
                u.y = pack_bf16(v[j + 2], v[j + 3]);
                u.z = pack_bf16(v[j + 4], v[j + 5]);
                u.w = pack_bf16(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(crow + j) = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + j < N) crow[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
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
static int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                         int box_rows) {
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
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const TcEpilogue& ep,
                     cudaStream_t stream) {
  using Cfg = TcCfg<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN, STAGES><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(ma, mb, M, N, K, ep);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

int linear_bf16_tc(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                   const LinearEpilogue& e, int tile_hint, cudaStream_t stream) {
  WF_REQUIRE(M > 0 && N > 0 && K > 0, "linear: empty problem M=%d N=%d K=%d", M, N, K);
  int bn = tile_hint;
  if (bn == 0) bn = (M <= 256) ? 32 : (N >= 256 ? 256 : 128);
  WF_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "linear: unsupported tile hint %d", tile_hint);
  CUtensorMap ma, mb;
  int rc = make_map_bf16(&ma, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_map_bf16(&mb, W, N, K, ldw, bn);
  if (rc) return rc;
  TcEpilogue ep;
  ep.C = e.C; ep.ldc = e.ldc; ep.bias = e.bias; ep.residual = e.residual; ep.ldr = e.ldr;
  ep.res_row_mod = e.res_row_mod; ep.gate = e.gate; ep.act = e.act; ep.out_f32 = e.out_f32;
  ep.c_off_ptr = e.c_off_ptr; ep.c_off_mul = e.c_off_mul;
  switch (bn) {
    case 32: return launch_tc<32, 8>(ma, mb, M, N, K, ep, stream);
    case 64: return launch_tc<64, 8>(ma, mb, M, N, K, ep, stream);
    case 128: return launch_tc<128, 6>(ma, mb, M, N, K, ep, stream);
    default: return launch_tc<256, 4>(ma, mb, M, N, K, ep, stream);
  }
}

}  // namespace wf
