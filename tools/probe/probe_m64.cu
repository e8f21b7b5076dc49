// Probe: where do the rows of a tcgen05.mma cta_group::1 M=64 accumulator land in tensor memory?
// A[r][0] = r + 1 (other k = 0), B[n][0] = 1  ->  D[r][n] = r + 1.  TMEM is zeroed first; every lane is dumped.
#include "../../whisper-flamingo_b200/csrc/common.cuh"
#include <vector>
using namespace wf;
__global__ void __launch_bounds__(128, 1) probe(float* out, int M) {
  __shared__ __align__(1024) uint8_t tiles[2 * 16384];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(tiles)[i] = 0;
  __syncthreads();
  // A tile: 128 rows x 64 cols bf16, K-major, 128B swizzle: (row, chunk) -> row * 128 + ((chunk ^ (row & 7)) << 4)
  if (tid < 128) {
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(tiles + tid * 128 + ((0 ^ (tid & 7)) << 4));
    a[0] = __float2bfloat16(static_cast<float>(tid + 1));
    if (tid < 32) {
      __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(tiles + 16384 + tid * 128 + ((0 ^ (tid & 7)) << 4));
      b[0] = __float2bfloat16(1.0f);
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<64>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  {  // zero the accumulator region (all 128 lanes x 32 columns)
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    tmem_st_32x32(tb + (static_cast<uint32_t>(warp * 32) << 16), z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, 32);
    umma_f16(tb, umma_desc_kmajor_sw128(smem_u32(tiles)), umma_desc_kmajor_sw128(smem_u32(tiles + 16384)), idesc, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tb + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tb);
}
int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  for (int M : {128, 64}) {
    cudaMemset(d, 0, 128 * 32 * 4);
    probe<<<1, 128>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * 32);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("M=%d (%s)\n", M, cudaGetErrorString(e));
    for (int l = 0; l < 128; ++l) {
      printf("lane %3d:", l);
      for (int c : {0, 1, 8, 15, 16, 17, 31}) printf(" c%d=%g", c, h[l * 32 + c]);
      printf("\n");
    }
  }
  return 0;
}
namespace wf { void set_error(const char*, ...) {} }
