// Probe: what does one narrow tcgen05.mma (cta_group::1, kind::f16, K = 16, both operands in shared memory) cost?
// One CTA, one (or two) issuing thread(s), REP back-to-back MMAs on zeroed operand tiles that rotate through a ring
// of shared-memory slots, one commit at the end; clock64 from the first issue to the completion of the commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_cost mma_cost.cu -lcuda
#include "../../whisper-flamingo_b200/csrc/common.cuh"
#include <cstdio>
#include <vector>
using namespace wf;

static constexpr int REP = 1024;
static constexpr int A_BYTES = 8 * 16384;   // 8 slots of [128 rows x 128 B]
static constexpr int B_BYTES = 2 * 32768;   // 2 slots of [256 rows x 128 B]

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// mode bit 0: A MN-major; bit 1: B MN-major; issuers: 1 or 2 (thread w = warp w lane 0, own accumulators)
__global__ void __launch_bounds__(128, 1) probe(long long* out, int M, int N, int mode, int issuers, int same_acc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (lane == 0 && warp < issuers) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + A_BYTES);
    uint32_t idesc = umma_idesc_bf16(M, N, (mode >> 1) & 1);
    if (mode & 1) idesc |= 1u << 15;
    const int reps = REP / issuers;
    // descriptors are loop-invariant registers and the loop is unrolled by 8: the issuing thread does next to no
    // integer work per MMA (a first version rebuilt them every iteration and measured its own ALU latency, 200 clk)
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ad[k] = (mode & 1) ? desc_mn(a0 + k * 2048, 16384) : umma_desc_kmajor_sw128(a0) + 2 * k;
      bd[k] = (mode & 2) ? desc_mn(b0 + k * 2048, 8192) : umma_desc_kmajor_sw128(b0) + 2 * k;
    }
    const uint32_t acc0 = tb + (issuers > 1 ? warp * 128 : 0);
    const uint32_t acc1 = acc0 + (same_acc ? 0 : (issuers > 1 ? 64 : 256));
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_f16((u & 4) ? acc1 : acc0, ad[u & 3], bd[u & 3], idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  const int smem = A_BYTES + B_BYTES + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int M, N, mode, issuers, same; const char* what; };
  std::vector<Cfg> cfgs = {
      {128, 32, 0, 1, 0, "M128 N32  A K-major  B K-major"},   {128, 64, 0, 1, 0, "M128 N64"},
      {128, 128, 0, 1, 0, "M128 N128"},                        {128, 256, 0, 1, 0, "M128 N256"},
      {64, 32, 0, 1, 0, "M64  N32"},                           {64, 64, 0, 1, 0, "M64  N64"},
      {64, 128, 0, 1, 0, "M64  N128"},                         {64, 256, 0, 1, 0, "M64  N256"},
      {128, 16, 0, 1, 0, "M128 N16"},                          {64, 16, 0, 1, 0, "M64  N16"},
      {64, 8, 0, 1, 0, "M64  N8"},
      {128, 32, 1, 1, 0, "M128 N32  A MN-major"},              {128, 32, 0, 1, 1, "M128 N32  one accumulator"},
      {128, 32, 0, 2, 0, "M128 N32  two issuing threads"},     {128, 32, 1, 2, 0, "M128 N32  A MN-major, two issuing threads"},
      {128, 32, 0, 3, 0, "M128 N32  three issuing threads"},   {128, 32, 0, 4, 0, "M128 N32  four issuing threads"},
      {128, 64, 0, 4, 0, "M128 N64  four issuing threads"},    {128, 128, 0, 2, 1, "M128 N128 two issuing threads"},
      {64, 256, 2, 1, 0, "M64  N256 B MN-major"},              {128, 256, 2, 1, 0, "M128 N256 B MN-major"},
      {64, 128, 2, 1, 0, "M64  N128 B MN-major"},
  };
  printf("tcgen05.mma kind::f16 K=16, %d back-to-back, clocks per MMA (issue loop | until commit completes)\n", REP);
  for (const Cfg& c : cfgs) {
    cudaMemset(d, 0, 8 * sizeof(long long));
    probe<<<1, 128, smem>>>(d, c.M, c.N, c.mode, c.issuers, c.same);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-45s CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long tot = 0;
    for (int w = 0; w < c.issuers; ++w) tot = h[2 * w + 1] > tot ? h[2 * w + 1] : tot;
    printf("%-45s issue %6.1f | total %6.1f clk/MMA   (operand bytes/MMA: A %d + B %d)\n", c.what,
           double(h[0]) / (REP / c.issuers), double(tot) / REP, c.M * 32, c.N * 32);
  }
  return 0;
}
