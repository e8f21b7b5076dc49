// KV-cached decode step kernels: split-K attention over cached K/V, greedy sampling with the logit
// filters fused, beam top-k, KV-row gather for beam reordering, device-side step bookkeeping.
//
// The reference recomputes the whole decoder (and the cross / x-attn K,V projections of all 1500 / T_x
// source frames) at every step (whisper/decoding.py:155-164, "disable kv cache"); these kernels give the
// same per-step result from cached K/V: softmax_fp32(q k^T / 8) v per head (model.py:93-108), then the
// filters + GreedyDecoder.update of decoding.py:432-442, 281-297.
//
// The step is HBM-bound: per (audio, head) the kernel streams 2 x len x 128 B of K/V exactly once, each
// thread pulling whole 128-byte rows (K) or 16-byte row slices (V) with many loads in flight; the G
// query rows of one audio (beams) share that single pass.  Everything position-dependent (current
// length, write offsets) is read from device memory so one captured CUDA graph serves every step.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace wf {

static constexpr int HD = 64;
static constexpr int DT = 128;       // threads per block == keys per tile
static constexpr int PART = HD + 2;  // floats per split partial: m, l, o[64]

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
  v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
}

template <typename T, int NQ>
__global__ void __launch_bounds__(DT)
attn_decode_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ kc, const T* __restrict__ vc,
                   long long ld_kv, long long kv_batch_stride, long long kv_head_stride, T* __restrict__ o,
                   long long ldo, int H, const int* __restrict__ len_ptr, int len_add, int len_const, int n_splits,
                   float* __restrict__ partials, const int* __restrict__ row_table, int table_ld, int q_group,
                   int q_first) {
  __shared__ float sq[NQ][HD];
  __shared__ float sp[NQ][DT];
  __shared__ float sred[NQ][DT / 32];
  __shared__ float sacc[DT / 32][NQ][HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  pdl_wait();
  const int kvb = blockIdx.x / H, h = blockIdx.x % H;
  const int split = blockIdx.y;
  const int len = len_ptr ? (*len_ptr + len_add) : len_const;
  const int n_tiles = (len + DT - 1) / DT;
  const int tiles_per_split = (n_tiles + n_splits - 1) / n_splits;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(n_tiles, tile_begin + tiles_per_split);

  for (int i = tid; i < NQ * HD; i += DT) {
    const int qi = i / HD, d = i % HD;
    sq[qi][d] = to_f32(q[(static_cast<long long>(kvb) * q_group + q_first + qi) * ldq + h * HD + d]) * 0.125f;
  }
  __syncthreads();

  // row_table (beam search): key j of cache entry kvb physically lives in entry row_table[kvb * table_ld + j] - the
  // hypotheses are re-ordered by rewriting this small table instead of moving K/V rows (decoding.py:173-180)
  const int* tbl = row_table ? row_table + static_cast<long long>(kvb) * table_ld : nullptr;
  const T* kbase = kc + h * kv_head_stride;
  const T* vbase = vc + h * kv_head_stride;
  const long long own = static_cast<long long>(kvb) * kv_batch_stride;
  const int dgrp = lane & 7, ksub = lane >> 3;

  float m_run[NQ], l_run[NQ], acc[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int key = tile * DT + tid;
    // ---- scores for this thread's key against the NQ queries
    float s[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) s[i] = 0.f;
    if (key < len) {
      const T* kr = kbase + (tbl ? static_cast<long long>(__ldg(tbl + key)) * kv_batch_stride : own) + key * ld_kv;
      float kv[8][8];
#pragma unroll
      for (int c = 0; c < 8; ++c) ld8(kr + c * 8, kv[c]);
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < NQ; ++i) s[i] = fmaf(sq[i][c * 8 + j], kv[c][j], s[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NQ; ++i) s[i] = -INFINITY;
    }
    // ---- tile max per query (block reduce)
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float wm = warp_max(s[i]);
      if (lane == 0) sred[i][warp] = wm;
    }
    __syncthreads();
    float corr[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float tm = sred[i][0];
#pragma unroll
      for (int w = 1; w < DT / 32; ++w) tm = fmaxf(tm, sred[i][w]);
      const float mn = fmaxf(m_run[i], tm);  // finite: every tile in range holds at least one valid key
      corr[i] = expf(m_run[i] - mn);
      m_run[i] = mn;
      const float p = expf(s[i] - mn);
      sp[i][tid] = p;
      l_run[i] = l_run[i] * corr[i] + p;  // per-thread partial sum of its own keys (reduced at the end)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] *= corr[i];
    }
    __syncthreads();
    // ---- acc += P V : warp w takes keys [32w, 32w+32) of the tile, 4 keys per instruction, 8 dims per lane
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int kt = warp * 32 + kk * 4 + ksub;
      const int vkey = tile * DT + kt;
      if (vkey < len) {
        float vv[8];
        ld8(vbase + (tbl ? static_cast<long long>(__ldg(tbl + vkey)) * kv_batch_stride : own) + vkey * ld_kv + dgrp * 8, vv);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float p = sp[i][kt];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(p, vv[j], acc[i][j]);
        }
      }
    }
    // sp / sred are rewritten only after the next tile's first barrier; reads above are complete by then
    __syncthreads();
  }

  // ---- reduce: l over all threads, acc over the 4 key sub-groups of a warp and the 4 warps
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const float wl = warp_sum(l_run[i]);
    if (lane == 0) sred[i][warp] = wl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[i][j];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      acc[i][j] = a;
    }
    if (ksub == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sacc[warp][i][dgrp * 8 + j] = acc[i][j];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < NQ * HD; idx += DT) {
    const int i = idx / HD, d = idx % HD;
    float a = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < DT / 32; ++w) { a += sacc[w][i][d]; l += sred[i][w]; }
    const long long row = static_cast<long long>(kvb) * q_group + q_first + i;
    if (n_splits == 1) {
      o[row * ldo + h * HD + d] = from_f32<T>(a / l);
    } else {
      float* pp = partials + ((row * H + h) * n_splits + split) * PART;
      pp[2 + d] = a;
      if (d == 0) { pp[0] = m_run[i]; pp[1] = l; }
    }
  }
}

// ---- growing self-attention cache (decode step: one query per hypothesis, a few dozen to a few hundred keys whose
// number lives in device memory): one WARP per (hypothesis, head).  The block-per-item kernel above spends its time in
// block barriers and launches 2560 CTAs of 128 threads for 128 x 20 items (1.08 waves of 148 x 16 resident CTAs: the
// tail wave doubles it); here an item's K and V rows are read in fully coalesced 512-byte pieces (8 lanes x 16 B per
// key, 4 keys per instruction), 32 keys per round with an online softmax, everything in registers and shuffles.
//   lane = (key group kg = lane / 8, dim group dg = lane % 8);  score of key 4 i + kg after round i lives in the 8 lanes
//   of group kg - lane (kg, dg) keeps the one of round i == dg, so each lane ends a round holding one key's score.
static constexpr int SW_WARPS = 8;
__global__ void __launch_bounds__(SW_WARPS * 32)
attn_decode_warp_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ kc,
                        const __nv_bfloat16* __restrict__ vc, long long ld_kv, long long kv_batch_stride,
                        long long kv_head_stride, __nv_bfloat16* __restrict__ o, long long ldo, int H, int items,
                        const int* __restrict__ len_ptr, int len_add, int len_const, const int* __restrict__ row_table,
                        int table_ld) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
  if (item >= items) return;
  const int row = item / H, h = item - row * H;
  const int len = len_ptr ? (*len_ptr + len_add) : len_const;
  const int kg = lane >> 3, dg = lane & 7;
  float qv[8];
  ld8(q + static_cast<long long>(row) * ldq + h * HD + dg * 8, qv);
#pragma unroll
  for (int j = 0; j < 8; ++j) qv[j] *= 0.125f;
  const int* tbl = row_table ? row_table + static_cast<long long>(row) * table_ld : nullptr;
  const long long own = static_cast<long long>(row) * kv_batch_stride + h * kv_head_stride + dg * 8;
  const long long hoff = static_cast<long long>(h) * kv_head_stride + dg * 8;
  float m_run = -INFINITY, l_run = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int base = 0; base < len; base += 32) {
    // physical cache entry of key base + lane (beam search re-orders hypotheses through this table)
    const int my_entry = (tbl && base + lane < len) ? __ldg(tbl + base + lane) : row;
    long long koff[8];
    float s_mine = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = base + 4 * i + kg;
      const int entry = tbl ? __shfl_sync(0xffffffffu, my_entry, 4 * i + kg) : row;
      koff[i] = (tbl ? static_cast<long long>(entry) * kv_batch_stride + hoff : own) + static_cast<long long>(key) * ld_kv;
      float part = 0.f;
      if (key < len) {
        float kv[8];
        ld8(kc + koff[i], kv);
#pragma unroll
        for (int j = 0; j < 8; ++j) part = fmaf(qv[j], kv[j], part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (i == dg && key < len) s_mine = part;
    }
    // online softmax over the 32 keys of the round (lane (kg, dg) holds key base + 4 dg + kg)
    const float m_new = fmaxf(m_run, warp_max(s_mine));
    const float corr = __expf(m_run - m_new);      // 0 on the first round
    const float p_mine = __expf(s_mine - m_new);   // 0 for keys beyond the length
    l_run = l_run * corr + warp_sum(p_mine);
    m_run = m_new;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = base + 4 * i + kg;
      const float p = __shfl_sync(0xffffffffu, p_mine, (kg << 3) | i);
      if (key < len) {
        float vv[8];
        ld8(vc + koff[i], vv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, vv[j], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  }
  if (kg == 0) {
    const float inv = 1.0f / l_run;
    uint4 u;
    u.x = pack_bf16(acc[0] * inv, acc[1] * inv);
    u.y = pack_bf16(acc[2] * inv, acc[3] * inv);
    u.z = pack_bf16(acc[4] * inv, acc[5] * inv);
    u.w = pack_bf16(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + static_cast<long long>(row) * ldo + h * HD + dg * 8) = u;
  }
}

// ---- head-major bf16 variant: the K (V) rows of one (audio, head) are contiguous 128-byte lines, so a tile of
// 128 keys is one contiguous 16 KB block.  Tiles are staged in shared memory by 1-D bulk copies (cp.async.bulk,
// mbarrier completion) three stages ahead of the math: the memory system always has 2 CTAs x 3 x 32 KB = 192 KB in
// flight per SM, independent of register pressure and of the softmax barriers.
static constexpr int HM_KEYS = 128;
static constexpr int HM_TILE_BYTES = HM_KEYS * HD * 2;           // 16 KB of K (or V)
static constexpr int HM_SMEM_BYTES = 3 * 2 * HM_TILE_BYTES + 128;

template <int NQ, int HM_STAGES = 3>
__global__ void __launch_bounds__(DT, 2)
attn_decode_hm_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ kc,
                      const __nv_bfloat16* __restrict__ vc, long long kv_batch_stride, long long kv_head_stride,
                      __nv_bfloat16* __restrict__ o, long long ldo, int H, const int* __restrict__ len_ptr,
                      int len_add, int len_const, int n_splits, float* __restrict__ partials, int q_group,
                      int q_first) {
  extern __shared__ uint8_t hm_raw[];
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hm_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t full_bar[HM_STAGES];
  __shared__ __align__(16) float sq[NQ][HD];  // read as float4
  __shared__ float sp[NQ][DT];
  __shared__ float sred[NQ][DT / 32];
  __shared__ float sacc[DT / 32][NQ][HD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  // A cache of static length (cross / x-attention) was written long before this step: its first tiles are requested
  // before the dependency wait and stream in while the previous kernel drains.  A growing cache (len_ptr) is not.
  if (len_ptr) pdl_wait();
  const int kvb = blockIdx.x / H, h = blockIdx.x % H;
  const int split = blockIdx.y;
  const int len = len_ptr ? (*len_ptr + len_add) : len_const;
  const int n_tiles = (len + HM_KEYS - 1) / HM_KEYS;
  const int tiles_per_split = (n_tiles + n_splits - 1) / n_splits;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(n_tiles, tile_begin + tiles_per_split);

  const __nv_bfloat16* kbase = kc + kvb * kv_batch_stride + h * kv_head_stride;
  const __nv_bfloat16* vbase = vc + kvb * kv_batch_stride + h * kv_head_stride;

  auto issue = [&](int tile, int stage) {
    const int rows = min(HM_KEYS, len - tile * HM_KEYS);
    const uint32_t bytes = static_cast<uint32_t>(rows) * HD * 2;
    uint8_t* dst = stage_base + stage * 2 * HM_TILE_BYTES;
    mbar_arrive_expect_tx(&full_bar[stage], 2 * bytes);
    bulk_load_1d(dst, kbase + static_cast<long long>(tile) * HM_KEYS * HD, bytes, &full_bar[stage]);
    bulk_load_1d(dst + HM_TILE_BYTES, vbase + static_cast<long long>(tile) * HM_KEYS * HD, bytes, &full_bar[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < HM_STAGES; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
    for (int s = 0; s < HM_STAGES; ++s)
      if (tile_begin + s < tile_end) issue(tile_begin + s, s);
  }
  if (!len_ptr) pdl_wait();  // q (and everything written below) depends on the previous kernel
  for (int i = tid; i < NQ * HD; i += DT) {
    const int qi = i / HD, d = i % HD;
    sq[qi][d] = __bfloat162float(q[(static_cast<long long>(kvb) * q_group + q_first + qi) * ldq + h * HD + d]) * 0.125f;
  }
  __syncthreads();

  const int dgrp = lane & 7, ksub = lane >> 3;
  float m_run[NQ], l_run[NQ], acc[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int rel = tile - tile_begin;
    const int stage = rel % HM_STAGES;
    const uint32_t parity = (rel / HM_STAGES) & 1;
    const uint8_t* ks = stage_base + stage * 2 * HM_TILE_BYTES;
    const uint8_t* vs = ks + HM_TILE_BYTES;
    mbar_wait(&full_bar[stage], parity);
    // ---- scores: thread = key; 16-byte chunks visited in a lane-rotated order so a quarter-warp hits 8 banks groups
    const int key = tile * HM_KEYS + tid;
    float s[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) s[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = (j + tid) & 7;
      const uint4 u = *reinterpret_cast<const uint4*>(ks + tid * 128 + cc * 16);
      const float kf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const float4 qa = *reinterpret_cast<const float4*>(&sq[i][cc * 8]);
        const float4 qb = *reinterpret_cast<const float4*>(&sq[i][cc * 8 + 4]);
        s[i] = fmaf(qa.x, kf[0], s[i]); s[i] = fmaf(qa.y, kf[1], s[i]); s[i] = fmaf(qa.z, kf[2], s[i]);
        s[i] = fmaf(qa.w, kf[3], s[i]); s[i] = fmaf(qb.x, kf[4], s[i]); s[i] = fmaf(qb.y, kf[5], s[i]);
        s[i] = fmaf(qb.z, kf[6], s[i]); s[i] = fmaf(qb.w, kf[7], s[i]);
      }
    }
    if (key >= len) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) s[i] = -INFINITY;  // stale smem rows of a partial tile
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float wm = warp_max(s[i]);
      if (lane == 0) sred[i][warp] = wm;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float tm = sred[i][0];
#pragma unroll
      for (int w = 1; w < DT / 32; ++w) tm = fmaxf(tm, sred[i][w]);
      const float mn = fmaxf(m_run[i], tm);
      const float corr = expf(m_run[i] - mn);
      m_run[i] = mn;
      const float p = expf(s[i] - mn);
      sp[i][tid] = p;
      l_run[i] = l_run[i] * corr + p;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] *= corr;
    }
    __syncthreads();
    // ---- acc += P V from shared memory: warp w takes keys [32w, 32w+32), 4 keys x 8 dim-groups per instruction
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int kt = warp * 32 + kk * 4 + ksub;
      if (tile * HM_KEYS + kt < len) {
        const uint4 u = *reinterpret_cast<const uint4*>(vs + kt * 128 + dgrp * 16);
        const float vf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float p = sp[i][kt];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(p, vf[j], acc[i][j]);
        }
      }
    }
    __syncthreads();  // every thread is done with this stage (and with sp / sred)
    if (tid == 0 && tile + HM_STAGES < tile_end) issue(tile + HM_STAGES, stage);
  }

#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const float wl = warp_sum(l_run[i]);
    if (lane == 0) sred[i][warp] = wl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[i][j];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      acc[i][j] = a;
    }
    if (ksub == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sacc[warp][i][dgrp * 8 + j] = acc[i][j];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < NQ * HD; idx += DT) {
    const int i = idx / HD, d = idx % HD;
    float a = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < DT / 32; ++w) { a += sacc[w][i][d]; l += sred[i][w]; }
    const long long row = static_cast<long long>(kvb) * q_group + q_first + i;
    if (n_splits == 1) {
      o[row * ldo + h * HD + d] = __float2bfloat16_rn(a / l);
    } else {
      float* pp = partials + ((row * H + h) * n_splits + split) * PART;
      pp[2 + d] = a;
      if (d == 0) { pp[0] = m_run[i]; pp[1] = l; }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(HD)
attn_decode_combine_kernel(const float* __restrict__ partials, T* __restrict__ o, long long ldo, int H, int n_splits,
                           int nq, int q_group, int q_first) {
  pdl_trigger();
  pdl_wait();
  const int item = blockIdx.x / H;  // (cache entry, query of this launch's chunk)
  const long long row = static_cast<long long>(item / nq) * q_group + q_first + item % nq;
  const int h = blockIdx.x % H, d = threadIdx.x;
  const float* pp = partials + (row * H + h) * n_splits * PART;
  float M = -INFINITY;
  for (int s = 0; s < n_splits; ++s) M = fmaxf(M, pp[s * PART]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float ms = pp[s * PART];
    const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
    num += w * pp[s * PART + 2 + d];
    den += w * pp[s * PART + 1];
  }
  o[row * ldo + h * HD + d] = from_f32<T>(num / den);
}

// ---- multi-query variant (beam search / best_of: the G hypotheses of a clip read the same cross / x-attention K,V).
// The thread-per-key kernels above spend G x 128 FMAs per key and drop to 1.9 TB/s at G = 5; here the G <= 8 query rows
// are the (zero-padded) 16-row A operand of warp-level tensor-core MMAs: S = Q K^T and O += P V as mma.sync m16n8k16,
// K/V tiles of 128 keys staged by TMA with the 128-byte swizzle (conflict-free ldmatrix), two stages = 64 KB per CTA,
// three CTAs per SM.  Each warp owns 32 keys of a tile with its own online softmax; the four partial results are
// merged once per (clip, head).
static constexpr int MQ_SMEM_BYTES = 2 * 2 * HM_TILE_BYTES + 2048 + 1024;  // 2 stages x (K + V) + Q tile + alignment

__device__ __forceinline__ const __nv_bfloat16* mq_sw(const uint8_t* tile, int row, int chunk) {
  return reinterpret_cast<const __nv_bfloat16*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int NQ>
__global__ void __launch_bounds__(DT, 3)
attn_decode_mq_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                      const __nv_bfloat16* __restrict__ q, long long ldq, int rows_per_batch, int rows_per_head,
                      __nv_bfloat16* __restrict__ o, long long ldo, int H, int len, int q_group, int q_first) {
  static_assert(NQ >= 1 && NQ <= 8, "queries per clip must fit the 8 real rows of the padded 16-row MMA operand");
  extern __shared__ uint8_t mq_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mq_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = base + 2 * 2 * HM_TILE_BYTES;  // [16][64] bf16, swizzled like the K/V tiles
  __shared__ __align__(8) uint64_t full_bar[2];
  __shared__ float comb_ml[4][8][2];
  float* comb_o = reinterpret_cast<float*>(base);  // [4 warps][8 rows][64] fp32, aliases stage 0 after the main loop
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const int kvb = blockIdx.x / H, h = blockIdx.x % H;
  const int n_tiles = (len + HM_KEYS - 1) / HM_KEYS;
  const int row0 = kvb * rows_per_batch + h * rows_per_head;  // first K (and V) row of this (clip, head) in the maps

  auto issue = [&](int tile, int stage) {
    uint8_t* dst = base + stage * 2 * HM_TILE_BYTES;
    mbar_arrive_expect_tx(&full_bar[stage], 2 * HM_TILE_BYTES);
    tma_load_2d(dst, &map_k, &full_bar[stage], 0, row0 + tile * HM_KEYS);
    tma_load_2d(dst + HM_TILE_BYTES, &map_v, &full_bar[stage], 0, row0 + tile * HM_KEYS);
  };
  if (tid == 0) {
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    mbar_fence_init();
    for (int s2 = 0; s2 < 2; ++s2)
      if (s2 < n_tiles) issue(s2, s2);  // the cache is static: requested before the dependency wait
  }
  pdl_wait();  // the queries come from the previous kernel
  {
    const int row = tid >> 3, chunk = tid & 7;  // 16 rows x 8 chunks of 16 B
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < NQ)
      v = *reinterpret_cast<const uint4*>(q + (static_cast<long long>(kvb) * q_group + q_first + row) * ldq + h * HD + chunk * 8);
    *reinterpret_cast<uint4*>(sQ + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
  }
  __syncthreads();
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qa[ks], mq_sw(sQ, lane & 15, ks * 2 + (lane >> 4)));

  constexpr float sl2 = 0.125f * 1.44269504088896340736f;  // 64^-0.5 * log2(e)
  float row_m = -INFINITY, row_l = 0.f;                    // this thread's query row = lane / 4 (rows 8..15 are padding)
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;

  for (int t = 0; t < n_tiles; ++t) {
    const int stage = t & 1;
    const uint8_t* ks_tile = base + stage * 2 * HM_TILE_BYTES;
    const uint8_t* vs_tile = ks_tile + HM_TILE_BYTES;
    mbar_wait(&full_bar[stage], (t >> 1) & 1);
    // ---- S = Q K^T for this warp's 32 keys: 4 n-tiles of 8 keys
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t kb[4];
        const int row = warp * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
        ldmatrix_x4(kb, mq_sw(ks_tile, row, ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s[np * 2], qa[ks], kb[0], kb[1]);
        mma_bf16_16816(s[np * 2 + 1], qa[ks], kb[2], kb[3]);
      }
    }
    const int kbase = t * HM_KEYS + warp * 32;
    if (kbase + 32 > len) {  // partial last tile: the rows past the end belong to the next head (or are zero fill)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (kbase + nt * 8 + (lane & 3) * 2 + j >= len) s[nt][j] = -INFINITY;
    }
    float mx = row_m;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mx = fmaxf(mx, fmaxf(s[nt][0], s[nt][1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_safe = (mx == -INFINITY) ? 0.f : mx;  // every key of this warp masked so far
    const float corr = ex2_approx((row_m - m_safe) * sl2);  // row_m = -inf -> 0
    const float msc = m_safe * sl2;
    row_m = mx;
    uint32_t pa[2][4];
    float psum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float p0 = ex2_approx(fmaf(s[nt][0], sl2, -msc));
      const float p1 = ex2_approx(fmaf(s[nt][1], sl2, -msc));
      psum += p0 + p1;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = 0u;  // padding rows 8..15
    }
    row_l = row_l * corr + psum;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      oacc[dt][0] *= corr;
      oacc[dt][1] *= corr;
    }
    // ---- O += P V : k = this warp's 32 keys (2 steps of 16), n = head_dim (8 tiles of 8)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int row = warp * 32 + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldmatrix_x4_trans(vb, mq_sw(vs_tile, row, dp * 2 + (lane >> 4)));
        mma_bf16_16816(oacc[dp * 2], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(oacc[dp * 2 + 1], pa[ks], vb[2], vb[3]);
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0 && t + 2 < n_tiles) issue(t + 2, stage);
  }

  // ---- merge the four warps' partial (m, l, O) and write the NQ output rows
  row_l += __shfl_xor_sync(0xffffffffu, row_l, 1);
  row_l += __shfl_xor_sync(0xffffffffu, row_l, 2);
  const int r = lane >> 2;
  if ((lane & 3) == 0) {
    comb_ml[warp][r][0] = row_m;
    comb_ml[warp][r][1] = row_l;
  }
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
    *reinterpret_cast<float2*>(comb_o + ((warp * 8 + r) * HD + dt * 8 + (lane & 3) * 2)) =
        make_float2(oacc[dt][0], oacc[dt][1]);
  __syncthreads();
  for (int idx = tid; idx < NQ * HD; idx += DT) {
    const int i = idx / HD, d = idx % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, comb_ml[w][i][0]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = comb_ml[w][i][0];
      const float sc = (mw == -INFINITY) ? 0.f : ex2_approx((mw - M) * sl2);
      num = fmaf(sc, comb_o[(w * 8 + i) * HD + d], num);
      den = fmaf(sc, comb_ml[w][i][1], den);
    }
    o[(static_cast<long long>(kvb) * q_group + q_first + i) * ldo + h * HD + d] = __float2bfloat16_rn(num / den);
  }
}

// ---- persistent variant for static-length caches (cross- and gated x-attention): ONE CTA per SM walks a strided list
// of (audio, head) items; a producer warp keeps PS_STAGES tiles of K/V (32 KB each, plus the item's query rows) in
// flight ACROSS item boundaries, four consumer warps run an independent online softmax each over their 32 keys of a
// tile (no block-wide barrier per tile) and merge once per item.  The CTA needs ~100 KB of shared memory, so a
// decode-step GEMM CTA of another sub-batch still fits on the SM next to it (SplitSession overlap).
static constexpr int PS_STAGES = 3;
static constexpr int PS_CONSUMERS = 128;
static constexpr int PS_THREADS = PS_CONSUMERS + 32;
template <int NQ>
struct PsSmem {
  static constexpr int Q_BYTES = NQ * HD * 2;                                   // bf16 query rows of an item
  static constexpr int STAGE_BYTES = 2 * HM_TILE_BYTES + ((Q_BYTES + 127) / 128) * 128;
  static constexpr int DYN_BYTES = PS_STAGES * STAGE_BYTES + 128;
};
__device__ __forceinline__ void ps_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NQ>
__global__ void __launch_bounds__(PS_THREADS, 1)
attn_decode_ps_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ kc,
                      const __nv_bfloat16* __restrict__ vc, long long kv_batch_stride, long long kv_head_stride,
                      __nv_bfloat16* __restrict__ o, long long ldo, int H, int len, int n_items, int q_group,
                      int q_first) {
  using L = PsSmem<NQ>;
  extern __shared__ uint8_t ps_raw[];
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ps_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t full_bar[PS_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[PS_STAGES];
  __shared__ __align__(16) float sq[NQ][HD];
  __shared__ float comb_acc[2][4][NQ][HD];
  __shared__ float comb_ml[2][4][NQ][2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const int n_tiles = (len + HM_KEYS - 1) / HM_KEYS;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < PS_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 4);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ------------------------------------------------------------ producer (one lane)
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      long long issued = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int kvb = item / H, h = item - kvb * H;
        const __nv_bfloat16* kbase = kc + kvb * kv_batch_stride + h * kv_head_stride;
        const __nv_bfloat16* vbase = vc + kvb * kv_batch_stride + h * kv_head_stride;
        for (int t = 0; t < n_tiles; ++t, ++issued) {
          if (issued >= PS_STAGES) mbar_wait(&empty_bar[stage], phase ^ 1);
          const int rows = min(HM_KEYS, len - t * HM_KEYS);
          const uint32_t bytes = static_cast<uint32_t>(rows) * HD * 2;
          uint8_t* dst = stage_base + stage * L::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], 2 * bytes + (t == 0 ? L::Q_BYTES : 0));
          bulk_load_1d_hint(dst, kbase + static_cast<long long>(t) * HM_KEYS * HD, bytes, &full_bar[stage], pol);
          bulk_load_1d_hint(dst + HM_TILE_BYTES, vbase + static_cast<long long>(t) * HM_KEYS * HD, bytes,
                            &full_bar[stage], pol);
          if (t == 0) {
            // the queries are written by the previous kernel: everything before this point streams a static cache
            if (!waited) { pdl_wait(); waited = true; }
#pragma unroll
            for (int i = 0; i < NQ; ++i)
              bulk_load_1d(dst + 2 * HM_TILE_BYTES + i * HD * 2,
                           q + (static_cast<long long>(kvb) * q_group + q_first + i) * ldq + h * HD, HD * 2,
                           &full_bar[stage]);
          }
          if (++stage == PS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (!waited) pdl_wait();
    }
    return;
  }

  // -------------------------------------------------------------- consumers: thread = key of the tile
  pdl_wait();  // o / anything written below must not race the previous kernel's readers
  const int dgrp = lane & 7, ksub = lane >> 3;
  constexpr float QSCALE = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e): softmax in base 2
  int stage = 0;
  uint32_t phase = 0;
  int parity = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, parity ^= 1) {
    const int kvb = item / H, h = item - kvb * H;
    float m_run[NQ], l_run[NQ], acc[NQ][8];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      m_run[i] = -INFINITY;
      l_run[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(&full_bar[stage], phase);
      const uint8_t* ks = stage_base + stage * L::STAGE_BYTES;
      const uint8_t* vs = ks + HM_TILE_BYTES;
      if (t == 0) {
        const __nv_bfloat16* qs = reinterpret_cast<const __nv_bfloat16*>(ks + 2 * HM_TILE_BYTES);
        for (int i = tid; i < NQ * HD; i += PS_CONSUMERS) sq[i / HD][i % HD] = __bfloat162float(qs[i]) * QSCALE;
        ps_bar();
      }
      const int key = t * HM_KEYS + tid;
      float s[NQ];
#pragma unroll
      for (int i = 0; i < NQ; ++i) s[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = (j + tid) & 7;
        const uint4 u = *reinterpret_cast<const uint4*>(ks + tid * 128 + cc * 16);
        const float kf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float4 qa = *reinterpret_cast<const float4*>(&sq[i][cc * 8]);
          const float4 qb = *reinterpret_cast<const float4*>(&sq[i][cc * 8 + 4]);
          s[i] = fmaf(qa.x, kf[0], s[i]); s[i] = fmaf(qa.y, kf[1], s[i]); s[i] = fmaf(qa.z, kf[2], s[i]);
          s[i] = fmaf(qa.w, kf[3], s[i]); s[i] = fmaf(qb.x, kf[4], s[i]); s[i] = fmaf(qb.y, kf[5], s[i]);
          s[i] = fmaf(qb.z, kf[6], s[i]); s[i] = fmaf(qb.w, kf[7], s[i]);
        }
      }
      float pr[NQ];
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        if (key >= len) s[i] = -INFINITY;  // stale rows of a partial tile
        const float mn = fmaxf(m_run[i], warp_max(s[i]));
        const bool dead = (mn == -INFINITY);
        const float corr = dead ? 1.f : ex2_approx(m_run[i] - mn);
        pr[i] = dead ? 0.f : ex2_approx(s[i] - mn);
        m_run[i] = mn;
        l_run[i] = l_run[i] * corr + pr[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] *= corr;
      }
      // acc += P V over this warp's 32 keys: 4 keys x 8 dim-groups per instruction, p by shuffle
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int kl = kk * 4 + ksub;
        const int kt = warp * 32 + kl;
        float pk[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) pk[i] = __shfl_sync(0xffffffffu, pr[i], kl);
        if (t * HM_KEYS + kt < len) {
          const uint4 u = *reinterpret_cast<const uint4*>(vs + kt * 128 + dgrp * 16);
          const float vf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
          for (int i = 0; i < NQ; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(pk[i], vf[j], acc[i][j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == PS_STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- merge the four warps' (m, l, acc) and write the item's output rows
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float wl = warp_sum(l_run[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = acc[i][j];
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        acc[i][j] = a;
      }
      if (ksub == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) comb_acc[parity][warp][i][dgrp * 8 + j] = acc[i][j];
      }
      if (lane == 0) {
        comb_ml[parity][warp][i][0] = m_run[i];
        comb_ml[parity][warp][i][1] = wl;
      }
    }
    ps_bar();
    for (int idx = tid; idx < NQ * HD; idx += PS_CONSUMERS) {
      const int i = idx / HD, d = idx % HD;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) M = fmaxf(M, comb_ml[parity][w][i][0]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mw = comb_ml[parity][w][i][0];
        const float sc = (mw == -INFINITY) ? 0.f : ex2_approx(mw - M);
        num = fmaf(sc, comb_acc[parity][w][i][d], num);
        den = fmaf(sc, comb_ml[parity][w][i][1], den);
      }
      o[(static_cast<long long>(kvb) * q_group + q_first + i) * ldo + h * HD + d] = __float2bfloat16_rn(num / den);
    }
  }
}

long long attention_decode_workspace_bytes(int R, int H) {
  return static_cast<long long>(R) * H * 32 /* max splits */ * PART * sizeof(float);
}

template <typename T, int NQ>
static int launch_decode_attn(const T* q, long long ldq, const T* kc, const T* vc, long long ld_kv,
                              long long kv_batch_stride, long long kv_head_stride, T* o, long long ldo, int R, int H,
                              const int* len_ptr,
                              int len_add, int len_max, float* ws, long long ws_bytes, const int* row_table,
                              int table_ld, int q_group, int q_first, cudaStream_t stream) {
  // R = rows of the whole problem (q_group per cache entry); this launch serves queries [q_first, q_first + NQ) of
  // every entry
  const int kvb = R / q_group;
  const int blocks = kvb * H;
  const int max_tiles = (len_max + DT - 1) / DT;
  // Granularity: a (audio, head) item streams up to 1500 x 256 B; with one CTA per item a 2560-item grid is
  // 1.08 waves of 148 x 16 resident CTAs and the tail wave costs ~45 %.  Split every item along the keys until
  // the grid is >= 8 waves (or one 128-key tile per CTA); partial (m, l, o) triples are merged by a second kernel.
  const long long want_ctas = 8LL * 16 * num_sms();
  int n_splits = static_cast<int>((want_ctas + blocks - 1) / blocks);
  if (const char* e = getenv("WF_DECODE_SPLITS")) { const int v = atoi(e); if (v > 0) n_splits = v; }
  if (n_splits > max_tiles) n_splits = max_tiles;
  if (n_splits > 32) n_splits = 32;
  if (n_splits < 1) n_splits = 1;
  {  // drop empty splits: ceil(tiles / splits) tiles each
    const int per = (max_tiles + n_splits - 1) / n_splits;
    n_splits = (max_tiles + per - 1) / per;
  }
  if (n_splits > 1)
    WF_REQUIRE(ws && ws_bytes >= static_cast<long long>(R) * H * n_splits * PART * (long long)sizeof(float),
               "attention_decode: workspace too small (need %lld bytes)",
               static_cast<long long>(R) * H * n_splits * PART * (long long)sizeof(float));
  if constexpr (sizeof(T) == 2 && NQ == 1) {
    // the growing self-attention cache of a decode step: warp-per-item kernel (WF_DECODE_WARP=0: the block-per-item one)
    static int warp_kernel = -1;
    if (warp_kernel < 0) {
      const char* e = getenv("WF_DECODE_WARP");
      warp_kernel = e ? atoi(e) : 1;
    }
    if (warp_kernel && len_ptr != nullptr && len_max <= 512 && q_group == 1 && ld_kv % 8 == 0 && ldq % 8 == 0 &&
        ldo % 8 == 0 && kv_batch_stride % 8 == 0 && kv_head_stride % 8 == 0 &&
        ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(kc) | reinterpret_cast<uintptr_t>(vc) |
          reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
      const int items = R * H;
      WF_CHECK_CUDA(launch_pdl(2, attn_decode_warp_kernel, dim3((items + SW_WARPS - 1) / SW_WARPS), dim3(SW_WARPS * 32), 0,
                               stream, (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kc, (const __nv_bfloat16*)vc,
                               ld_kv, kv_batch_stride, kv_head_stride, (__nv_bfloat16*)o, ldo, H, items, len_ptr, len_add,
                               len_max, row_table, table_ld));
      count_launch();
      return WF_OK;
    }
  }
  if constexpr (sizeof(T) == 2) if (!row_table) {
    // head-major cache + long key range: bulk-copy pipelined kernel, one CTA streams a whole (audio, head) item
    // (measured: 5.8 TB/s unsplit vs 4.0 TB/s when the item is cut into 12 one-tile CTAs - the pipeline needs depth)
    if (NQ >= 2 && NQ <= 8 && ld_kv == HD && !len_ptr && len_max >= HM_KEYS && kv_batch_stride % HD == 0 &&
        kv_head_stride % HD == 0 && getenv("WF_DECODE_NO_MQ") == nullptr) {
      constexpr int NQC = (NQ >= 2 && NQ <= 8) ? NQ : 2;  // the branch is only taken for 2..8
      const long long rpb = kv_batch_stride / HD, rph = kv_head_stride / HD;
      CUtensorMap mk, mv;
      int rc = make_map_bf16(&mk, kc, (kvb - 1) * rpb + (H - 1) * rph + len_max, HD, HD, HM_KEYS);
      if (rc) return rc;
      rc = make_map_bf16(&mv, vc, (kvb - 1) * rpb + (H - 1) * rph + len_max, HD, HD, HM_KEYS);
      if (rc) return rc;
      static PerDeviceOnce mq_configured;  // function attributes are per device
      if (mq_configured.first_use()) {
        WF_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_mq_kernel<NQC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           MQ_SMEM_BYTES));
      }
      WF_CHECK_CUDA(launch_pdl(2, attn_decode_mq_kernel<NQC>, dim3(blocks), dim3(DT), MQ_SMEM_BYTES, stream, mk, mv,
                               (const __nv_bfloat16*)q, ldq, static_cast<int>(rpb), static_cast<int>(rph),
                               (__nv_bfloat16*)o, ldo, H, len_max, q_group, q_first));
      count_launch();
      return WF_OK;
    }
    static int persist = -1;
    if (persist < 0) {
      const char* e = getenv("WF_DECODE_PERSIST");
      persist = e ? atoi(e) : 0;  // measured slower than the per-item kernel (4.1 vs 6.0 TB/s): opt-in
    }
    if (persist && ld_kv == HD && !len_ptr && len_max >= 256 && blocks >= 2 * num_sms()) {
      using L = PsSmem<NQ>;
      static PerDeviceOnce ps_configured;  // function attributes are per device
      if (ps_configured.first_use()) {
        WF_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_ps_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           L::DYN_BYTES));
      }
      WF_CHECK_CUDA(launch_pdl(2, attn_decode_ps_kernel<NQ>, dim3(num_sms()), dim3(PS_THREADS), L::DYN_BYTES, stream, q,
                               ldq, kc, vc, kv_batch_stride, kv_head_stride, o, ldo, H, len_max, blocks, q_group, q_first));
      count_launch();
      return WF_OK;
    }
    if (ld_kv == HD && len_max >= 512 && getenv("WF_DECODE_NO_BULK") == nullptr) {
      int hs = 1;
      if (const char* e = getenv("WF_DECODE_SPLITS")) { const int v = atoi(e); if (v > 0) hs = v < max_tiles ? v : max_tiles; }
      if (hs > 1)
        WF_REQUIRE(ws && ws_bytes >= static_cast<long long>(R) * H * hs * PART * (long long)sizeof(float),
                   "attention_decode: workspace too small");
      static int pad = -1;  // WF_ATTN_SMEM_PAD: extra dynamic shared memory (bytes) to cap the CTAs per SM (experiments)
      if (pad < 0) {
        const char* e = getenv("WF_ATTN_SMEM_PAD");
        pad = e ? atoi(e) : 0;
      }
      static PerDeviceOnce configured;  // function attributes are per device
      if (configured.first_use()) {
        WF_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_hm_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           HM_SMEM_BYTES + pad));
      }
      dim3 hgrid(blocks, hs);
      // Two stages (64 KB per CTA) let THREE CTAs share an SM instead of two: the softmax barriers and the prologue /
      // epilogue of one CTA hide behind the other two.  Measured on B200, large-v2 decode loop: 852 -> 792 ms.
      static int two = -1;
      if (two < 0) {
        const char* e = getenv("WF_ATTN_STAGES");
        two = (e && atoi(e) == 3) ? 0 : 1;
      }
      if (two) {
        constexpr int SM2 = 2 * 2 * HM_TILE_BYTES + 128;
        static PerDeviceOnce c2;  // function attributes are per device
        if (c2.first_use()) {
          WF_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_hm_kernel<NQ, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM2));
        }
        WF_CHECK_CUDA(launch_pdl(2, attn_decode_hm_kernel<NQ, 2>, hgrid, dim3(DT), SM2, stream, q, ldq, kc, vc,
                                 kv_batch_stride, kv_head_stride, o, ldo, H, len_ptr, len_add, len_max, hs, ws, q_group,
                                 q_first));
      } else {
        WF_CHECK_CUDA(launch_pdl(2, attn_decode_hm_kernel<NQ>, hgrid, dim3(DT), HM_SMEM_BYTES + pad, stream, q, ldq, kc, vc,
                                 kv_batch_stride, kv_head_stride, o, ldo, H, len_ptr, len_add, len_max, hs, ws, q_group,
                                 q_first));
      }
      count_launch();
      if (hs > 1) {
        WF_CHECK_CUDA(launch_pdl(2, attn_decode_combine_kernel<T>, dim3(kvb * NQ * H), dim3(HD), 0, stream,
                                 (const float*)ws, o, ldo, H, hs, NQ, q_group, q_first));
        count_launch();
      }
      return WF_OK;
    }
  }
  dim3 grid(blocks, n_splits);
  WF_CHECK_CUDA(launch_pdl(2, attn_decode_kernel<T, NQ>, grid, dim3(DT), 0, stream, q, ldq, kc, vc, ld_kv,
                           kv_batch_stride, kv_head_stride, o, ldo, H, len_ptr, len_add, len_max, n_splits, ws,
                           row_table, table_ld, q_group, q_first));
  count_launch();
  if (n_splits > 1) {
    WF_CHECK_CUDA(launch_pdl(2, attn_decode_combine_kernel<T>, dim3(kvb * NQ * H), dim3(HD), 0, stream, (const float*)ws,
                             o, ldo, H, n_splits, NQ, q_group, q_first));
    count_launch();
  }
  return WF_OK;
}

template <typename T>
static int dispatch_decode_attn(const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                                long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R,
                                int G, int H,
                                const int* len_ptr, int len_add, int len_max, void* ws, long long ws_bytes,
                                const int* row_table, int table_ld, cudaStream_t stream) {
  // Any G: the hypotheses of a cache entry are served in chunks of at most 8 queries (one launch per chunk, each
  // streaming the entry's K/V once); beam_size / best_of up to 8 is a single pass.
#define WF_DA(NQ)                                                                                              \
  case NQ:                                                                                                     \
    rc = launch_decode_attn<T, NQ>((const T*)q, ldq, (const T*)kc, (const T*)vc, ld_kv, kv_batch_stride,       \
                                   kv_head_stride, (T*)o, ldo, R, H, len_ptr, len_add, len_max, (float*)ws,    \
                                   ws_bytes, row_table, table_ld, G, first, stream);                           \
    break
  for (int first = 0; first < G; first += 8) {
    int rc = WF_OK;
    switch (G - first < 8 ? G - first : 8) {
      WF_DA(1); WF_DA(2); WF_DA(3); WF_DA(4); WF_DA(5); WF_DA(6); WF_DA(7); WF_DA(8);
    }
    if (rc) return rc;
  }
  return WF_OK;
#undef WF_DA
}

int attention_decode(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                     long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R, int G, int H,
                     const int* len_ptr, int len_add, int len_const, void* workspace, long long workspace_bytes,
                     const int* row_table, int table_ld, cudaStream_t stream) {
  WF_REQUIRE(!row_table || (G == 1 && table_ld >= len_const), "attention_decode: a row table needs G == 1 and table_ld >= max length");
  WF_REQUIRE(R > 0 && G > 0 && H > 0 && R % G == 0, "attention_decode: bad shape R=%d G=%d H=%d", R, G, H);
  WF_REQUIRE(len_const > 0, "attention_decode: len (or max len) must be positive");
  const int al = dtype == WF_BF16 ? 8 : 4;
  WF_REQUIRE(ld_kv % al == 0 && kv_batch_stride % al == 0 && kv_head_stride % al == 0,
             "attention_decode: K/V strides must keep 16-byte alignment");
  WF_REQUIRE((reinterpret_cast<uintptr_t>(kc) & 15) == 0 && (reinterpret_cast<uintptr_t>(vc) & 15) == 0,
             "attention_decode: K/V base must be 16-byte aligned");
  if (dtype == WF_F32)
    return dispatch_decode_attn<float>(q, ldq, kc, vc, ld_kv, kv_batch_stride, kv_head_stride, o, ldo, R, G, H,
                                       len_ptr, len_add, len_const, workspace, workspace_bytes, row_table, table_ld, stream);
  if (dtype == WF_BF16)
    return dispatch_decode_attn<__nv_bfloat16>(q, ldq, kc, vc, ld_kv, kv_batch_stride, kv_head_stride, o, ldo, R, G,
                                               H, len_ptr, len_add, len_const, workspace, workspace_bytes, row_table,
                                               table_ld, stream);
  WF_REQUIRE(false, "attention_decode: bad dtype %d", dtype);
}

// ============================================================================ sampling
struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  // larger value wins; ties -> lower index (torch.argmax convention)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  ArgMax r = sh[0];
  for (int w = 1; w < nw; ++w) r = better(r, sh[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float x, float* sh) {
  x = warp_sum(x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < nw; ++w) r += sh[w];
  return r;
}

// ---- logit filters as a per-row predicate (decoding.py:427-509).  Every rule of ApplyTimestampRules
// masks an index range, so a row's filter state is four integers computed once per row.
struct RowMask {
  int text_end;   // mask [0, text_end)
  int ts_upto;    // mask [timestamp_begin, ts_upto)
  int ts_from;    // mask [ts_from, V)
  int no_ts;      // single masked id or -1
  int tb;         // timestamp_begin (V when the rules are off)
  bool first;     // first sampled token -> SuppressBlank mask applies
};
__device__ __forceinline__ bool is_masked(const RowMask& m, int i, const uint8_t* __restrict__ suppress,
                                          const uint8_t* __restrict__ suppress_first) {
  if (suppress[i]) return true;
  if (m.first && suppress_first && suppress_first[i]) return true;
  if (i == m.no_ts || i < m.text_end || i >= m.ts_from) return true;
  return i >= m.tb && i < m.ts_upto;
}
// row = token history of this hypothesis, cur_len tokens of which the first n_init are the prompt
__device__ RowMask make_row_mask(const int* row, int n_init, int cur_len, int V, int eot, int tb, int no_ts,
                                 int max_initial_ts) {
  RowMask m;
  m.text_end = 0; m.ts_upto = 0; m.ts_from = V; m.no_ts = -1; m.tb = V;
  m.first = (cur_len == n_init);
  if (tb < 0) return m;
  m.tb = tb; m.ts_upto = tb; m.no_ts = no_ts;
  const int n_s = cur_len - n_init;
  const bool last_ts = n_s >= 1 && row[cur_len - 1] >= tb;
  const bool pen_ts = n_s < 2 || row[cur_len - 2] >= tb;
  if (last_ts) {
    if (pen_ts) m.ts_from = tb;   // has to be non-timestamp
    else m.text_end = eot;        // cannot be normal text tokens
  }
  int ts_val = -1;
  for (int j = cur_len - 1; j >= n_init; --j)
    if (row[j] >= tb) { ts_val = row[j]; break; }
  if (ts_val >= 0) m.ts_upto = (last_ts && !pen_ts) ? ts_val : ts_val + 1;  // timestamps must not decrease
  if (m.first) {
    m.text_end = tb;  // the first sampled token must be a timestamp
    if (max_initial_ts >= 0) m.ts_from = min(m.ts_from, tb + max_initial_ts + 1);
  }
  return m;
}
// "if the probability mass over timestamps exceeds every single text token, sample a timestamp" (decoding.py:501-509)
__device__ void apply_timestamp_mass_rule(RowMask& m, const float* lg, int V, const uint8_t* suppress,
                                          const uint8_t* suppress_first, ArgMax* sh_am, float* sh_f) {
  if (m.tb >= V) return;
  ArgMax text{-INFINITY, 0x7fffffff}, ts{-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    if (is_masked(m, i, suppress, suppress_first)) continue;
    if (i < m.tb) text = better(text, ArgMax{lg[i], i});
    else ts = better(ts, ArgMax{lg[i], i});
  }
  text = block_argmax(text, sh_am);
  ts = block_argmax(ts, sh_am);
  if (ts.v == -INFINITY) return;  // no timestamp allowed at all
  float se = 0.f;
  for (int i = m.tb + threadIdx.x; i < V; i += blockDim.x)
    if (!is_masked(m, i, suppress, suppress_first)) se += expf(lg[i] - ts.v);
  se = block_sum(se, sh_f);
  // logsumexp(ts) > max(text)  (both sides share the softmax normaliser)
  if (ts.v + logf(se) > text.v) m.text_end = m.tb;
}

// counter-based uniform in (0,1): splitmix64 finaliser over (seed, row, position, token id)
__device__ __forceinline__ float uniform01(unsigned long long seed, int row, int pos, int tok) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (static_cast<unsigned long long>(row) * 1000003ull + pos) +
                         0xD1B54A32D192ED03ull * static_cast<unsigned long long>(tok + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  // 23 random bits + 0.5 is exact in fp32 (24 significant bits), so u lies strictly inside (0, 1): [2^-24, 1 - 2^-24].
  // (24 bits + 0.5 rounds 0xFFFFFF + 0.5 up to 2^24, u == 1, Gumbel = +inf: that token would win whatever its logit.)
  return (static_cast<float>(z >> 41) + 0.5f) * (1.0f / 8388608.0f);
}

// test hook: min / max of n consecutive draws (token ids first + i of row `row`, position `pos`), float-as-int atomics
__global__ void uniform_range_kernel(unsigned long long seed, int row, int pos, long long first, long long n,
                                     unsigned int* out) {
  float lo = 2.f, hi = -1.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long t = first + i;
    const float u = uniform01(seed + static_cast<unsigned long long>(t >> 20), row, pos, static_cast<int>(t & 0xFFFFF));
    lo = fminf(lo, u);
    hi = fmaxf(hi, u);
  }
  lo = -warp_max(-lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {  // u > 0, so the unsigned integer order of the bit patterns is the float order
    atomicMin(&out[0], __float_as_uint(lo));
    atomicMax(&out[1], __float_as_uint(hi));
  }
}
int debug_uniform_range(unsigned long long seed, long long n, float* out_min_max, cudaStream_t stream) {
  WF_REQUIRE(n > 0 && out_min_max, "debug_uniform_range: bad arguments");
  const unsigned int init[2] = {0x7f800000u, 0u};
  WF_CHECK_CUDA(cudaMemcpyAsync(out_min_max, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  uniform_range_kernel<<<num_sms() * 8, 256, 0, stream>>>(seed, 3, 7, 0, n, reinterpret_cast<unsigned int*>(out_min_max));
  WF_CHECK_LAUNCH();
  return WF_OK;
}

// state: [0]=t (position of the token fed this step) [1]=n_init [2]=all_done [3]=#rows at EOT this step [4]=sot_index
//        [5],[6]=low/high word of a per-call RNG seed (temperature > 0), XORed with SampleArgs.seed: the seed changes
//        from call to call without re-capturing the CUDA graph the launch is part of
// One CTA of 1024 threads per row: the two passes over the 51865 fp32 logits of a row are latency-bound, so the row is
// spread over 32 warps (256 threads measured 110 us per step at R = 128, i.e. 0.5 TB/s).
__global__ void __launch_bounds__(1024) sample_greedy_kernel(SampleArgs a) {
  __shared__ ArgMax sh_am[32];
  __shared__ float sh_f[32];
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const float* lg = a.logits + r * a.ld;
  const int t = a.state[0], n_init = a.state[1], sot_index = a.state[4];

  if (a.no_speech >= 0 && t == sot_index) {
    // no_speech_prob = softmax(raw logits at the SOT position)[no_speech]   (decoding.py:697-701)
    ArgMax am{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < a.V; i += blockDim.x) am = better(am, ArgMax{lg[i], i});
    am = block_argmax(am, sh_am);
    float se = 0.f;
    for (int i = threadIdx.x; i < a.V; i += blockDim.x) se += expf(lg[i] - am.v);
    se = block_sum(se, sh_f);
    if (threadIdx.x == 0) a.no_speech_prob[r] = expf(lg[a.no_speech] - am.v) / se;
  }
  if (t + 1 < n_init) return;  // still feeding the forced initial tokens

  int* row = a.tokens + static_cast<long long>(r) * a.T_cap;
  RowMask m = make_row_mask(row, n_init, t + 1, a.V, a.eot, a.timestamp_begin, a.no_timestamps, a.max_initial_ts);
  apply_timestamp_mass_rule(m, lg, a.V, a.suppress, a.suppress_first, sh_am, sh_f);

  ArgMax am{-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < a.V; i += blockDim.x)
    if (!is_masked(m, i, a.suppress, a.suppress_first)) am = better(am, ArgMax{lg[i], i});
  am = block_argmax(am, sh_am);
  float se = 0.f;
  for (int i = threadIdx.x; i < a.V; i += blockDim.x)
    if (!is_masked(m, i, a.suppress, a.suppress_first)) se += expf(lg[i] - am.v);
  se = block_sum(se, sh_f);
  int chosen = am.i;
  float chosen_logit = am.v;
  if (a.temperature > 0.f) {
    // Gumbel-max: argmax_i (x_i / T + G_i), G_i = -log(-log u_i)  ~  Categorical(softmax(x / T))
    const float inv_t = 1.0f / a.temperature;
    const unsigned long long seed = a.seed ^ (static_cast<unsigned long long>(static_cast<unsigned>(a.state[5])) |
                                              (static_cast<unsigned long long>(static_cast<unsigned>(a.state[6])) << 32));
    ArgMax gm{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < a.V; i += blockDim.x)
      if (!is_masked(m, i, a.suppress, a.suppress_first)) {
        const float u = uniform01(seed, r, t, i);
        gm = better(gm, ArgMax{lg[i] * inv_t - logf(-logf(u)), i});
      }
    gm = block_argmax(gm, sh_am);
    chosen = gm.i;
    chosen_logit = lg[chosen];
  }
  if (threadIdx.x == 0) {
    const int prev = row[t];
    int next = chosen;
    if (prev == a.eot) {
      next = a.eot;  // finished rows keep emitting EOT and stop accumulating (decoding.py:291-293)
    } else {
      // log_softmax(filtered logits)[next] = x_next - max - log sum exp(x - max)   (un-tempered, decoding.py:289-291)
      a.sum_logprobs[r] += (chosen_logit - am.v) - logf(se);
    }
    row[t + 1] = next;
    if (next == a.eot) atomicAdd(&a.state[3], 1);
  }
}

int sample_greedy(const SampleArgs& a, cudaStream_t stream) {
  WF_REQUIRE(a.R > 0 && a.V > 0 && a.logits && a.tokens && a.state && a.suppress, "sample_greedy: bad arguments");
  WF_CHECK_CUDA(launch_pdl(3, sample_greedy_kernel, dim3(a.R), dim3(1024), 0, stream, a));
  count_launch();
  return WF_OK;
}

__global__ void step_advance_kernel(int* state, int R) {
  pdl_trigger();
  pdl_wait();
  const int t = state[0], n_init = state[1];
  if (state[2] != 0) {   // finished: steps replayed before the host has polled the flag change nothing
    state[3] = 0;
    return;
  }
  if (t + 1 >= n_init && state[3] == R) state[2] = 1;
  state[3] = 0;
  state[0] = t + 1;
}
int step_advance(int* state, int R, cudaStream_t stream) {
  WF_CHECK_CUDA(launch_pdl(3, step_advance_kernel, dim3(1), dim3(1), 0, stream, state, R));
  count_launch();
  return WF_OK;
}

// log_softmax over the filtered logits + top-k per row (BeamSearchDecoder.update, decoding.py:337-347)
__global__ void __launch_bounds__(256) topk_logprobs_kernel(TopkArgs a) {
  __shared__ ArgMax sh_am[8];
  __shared__ float sh_f[8];
  __shared__ int picked[32];
  const int r = blockIdx.x;
  const float* lg = a.logits + r * a.ld;
  if (a.state != nullptr) {
    // graph-replayable form: the position lives in device memory (state[0] = position of the token fed this step)
    if (a.state[2] != 0) return;   // search finished (the host polls the flag every few steps)
    const int t = a.state[0], sot_index = a.state[4];
    a.n_init = a.state[1];
    a.cur_len = t + 1;
    if (a.no_speech_prob != nullptr && a.no_speech >= 0 && t == sot_index) {
      // no_speech_prob = softmax(raw logits at the SOT position)[no_speech]   (decoding.py:697-701)
      ArgMax am{-INFINITY, 0x7fffffff};
      for (int i = threadIdx.x; i < a.V; i += blockDim.x) am = better(am, ArgMax{lg[i], i});
      am = block_argmax(am, sh_am);
      float se = 0.f;
      for (int i = threadIdx.x; i < a.V; i += blockDim.x) se += expf(lg[i] - am.v);
      se = block_sum(se, sh_f);
      if (threadIdx.x == 0) a.no_speech_prob[r] = expf(lg[a.no_speech] - am.v) / se;
      __syncthreads();
    }
    if (t + 1 < a.n_init) return;  // still feeding the forced initial tokens
  }
  RowMask m;
  if (a.tokens) {
    m = make_row_mask(a.tokens + static_cast<long long>(r) * a.T_cap, a.n_init, a.cur_len, a.V, a.eot,
                      a.timestamp_begin, a.no_timestamps, a.max_initial_ts);
  } else {
    m.text_end = 0; m.ts_upto = 0; m.ts_from = a.V; m.no_ts = -1; m.tb = a.V; m.first = (a.cur_len == a.n_init);
  }
  apply_timestamp_mass_rule(m, lg, a.V, a.suppress, a.suppress_first, sh_am, sh_f);
  float lse = 0.f, vmax = 0.f;
  for (int it = 0; it < a.k; ++it) {
    ArgMax am{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < a.V; i += blockDim.x) {
      bool skip = is_masked(m, i, a.suppress, a.suppress_first);
      for (int p = 0; p < it && !skip; ++p) skip = (picked[p] == i);
      if (!skip) am = better(am, ArgMax{lg[i], i});
    }
    am = block_argmax(am, sh_am);
    if (it == 0) {
      vmax = am.v;
      float se = 0.f;
      for (int i = threadIdx.x; i < a.V; i += blockDim.x)
        if (!is_masked(m, i, a.suppress, a.suppress_first)) se += expf(lg[i] - vmax);
      lse = logf(block_sum(se, sh_f));
    }
    if (threadIdx.x == 0) {
      picked[it] = am.i;
      a.out_vals[r * a.k + it] = (am.v - vmax) - lse;  // -inf when fewer than k tokens are allowed
      a.out_idx[r * a.k + it] = am.i == 0x7fffffff ? 0 : am.i;
    }
    __syncthreads();
  }
}
int topk_logprobs(const TopkArgs& a, cudaStream_t stream) {
  WF_REQUIRE(a.R > 0 && a.V > 0 && a.k > 0 && a.k <= 32, "topk: bad arguments (k=%d)", a.k);
  WF_REQUIRE(a.logits && a.suppress && a.out_vals && a.out_idx, "topk: null buffer");
  topk_logprobs_kernel<<<a.R, 256, 0, stream>>>(a);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

// One step of BeamSearchDecoder.update (decoding.py:337-386) on the device, one CTA per audio.
// Candidates = the top-(G + 1) continuations of each of the G hypotheses; identical hypotheses (all beams share the
// prompt at the first step; `hyp_id` equal <=> token sequences equal) collapse into one dictionary entry that keeps the
// position of its first insertion and the value of its last, as the reference's tuple-keyed dict does; entries are
// visited by descending cumulative log-probability (stable), EOT continuations go to the finished list (at most
// max_candidates per audio), the first G others become the next hypotheses.  The token histories and the per-position
// row table of the self-attention cache (decoding.py:173-180 rearrange_kv_cache) are permuted through scratch rows.
static constexpr int BEAM_MAX_CAND = 32 * 33;
__global__ void __launch_bounds__(128) beam_update_kernel(BeamArgs a) {
  __shared__ float c_score[BEAM_MAX_CAND];
  __shared__ int c_tok[BEAM_MAX_CAND];
  __shared__ short c_src[BEAM_MAX_CAND];      // hypothesis (0 .. G-1) the candidate extends
  __shared__ short c_order[BEAM_MAX_CAND];    // candidate index by rank (alive entries only)
  __shared__ int keep[32], fin_src[32], fin_slot[32];
  __shared__ float fin_sc[32];
  __shared__ int n_alive, n_keep, n_done;
  const int au = blockIdx.x, G = a.G, k = G + 1, n = G * k;
  const int t = a.state[0], n_init = a.state[1];
  if (t + 1 < n_init || a.state[2] != 0) return;   // feeding the forced initial tokens / search finished
  const int L = t + 1;                        // tokens per hypothesis so far
  const int row0 = au * G;
  if (threadIdx.x == 0) { n_alive = 0; n_keep = 0; n_done = 0; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int j = i / k, r = row0 + j;
    c_score[i] = a.sum_logprobs[r] + a.vals[r * k + (i - j * k)];    // fp32 add, as the reference
    c_tok[i] = a.idx[r * k + (i - j * k)];
    c_src[i] = static_cast<short>(j);
  }
  __syncthreads();
  // dictionary semantics: an entry lives at its first insertion and carries the value of its last
  float my_score[(BEAM_MAX_CAND + 127) / 128];
  short my_src[(BEAM_MAX_CAND + 127) / 128];
  bool my_alive[(BEAM_MAX_CAND + 127) / 128];
#pragma unroll 1
  for (int q = 0, i = threadIdx.x; i < n; i += blockDim.x, ++q) {
    const int hid = a.hyp_id[row0 + c_src[i]], tok = c_tok[i];
    bool first = true;
    int last = i;
    for (int i2 = 0; i2 < n; ++i2)
      if (i2 != i && c_tok[i2] == tok && a.hyp_id[row0 + c_src[i2]] == hid) {
        if (i2 < i) first = false;
        if (i2 > last) last = i2;
      }
    my_alive[q] = first;
    my_score[q] = c_score[last];
    my_src[q] = c_src[last];
  }
  __syncthreads();
#pragma unroll 1
  for (int q = 0, i = threadIdx.x; i < n; i += blockDim.x, ++q) {
    c_score[i] = my_alive[q] ? my_score[q] : NAN;     // NaN marks a collapsed duplicate
    c_src[i] = my_src[q];
  }
  __syncthreads();
  // rank among the live entries: descending score, ties by insertion order (Python's sort is stable)
#pragma unroll 1
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = c_score[i];
    if (v != v) continue;
    int rank = 0;
    for (int i2 = 0; i2 < n; ++i2) {
      const float w = c_score[i2];
      if (w != w) continue;
      if (w > v || (w == v && i2 < i)) ++rank;
    }
    c_order[rank] = static_cast<short>(i);
    atomicAdd(&n_alive, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int kept = 0, nd = 0, nf = a.n_fin[au];
    for (int p = 0; p < n_alive && kept < G; ++p) {
      const int i = c_order[p];
      if (c_tok[i] == a.eot) {
        if (nd < 32) { fin_src[nd] = i; ++nd; }
      } else {
        keep[kept++] = i;
      }
    }
    // finished sequences of this step in the order visited (= descending score), while there is room (:374-378)
    int nrec = 0;
    for (int q = 0; q < nd && nf < a.max_candidates; ++q) {
      const int i = fin_src[q];
      fin_slot[nrec] = nf;
      fin_sc[nrec] = c_score[i];
      fin_src[nrec] = row0 + c_src[i];
      ++nrec;
      ++nf;
    }
    a.n_fin[au] = nf;
    n_keep = kept;
    n_done = nrec;
    if (nf >= a.max_candidates) atomicAdd(&a.state[3], 1);
  }
  __syncthreads();
  const long long ld = a.T_cap;
  // finished: tokens of the parent + EOT
  for (int q = 0; q < n_done; ++q) {
    int* dst = a.fin_tokens + (static_cast<long long>(au) * a.max_candidates + fin_slot[q]) * ld;
    const int* src = a.tokens + static_cast<long long>(fin_src[q]) * ld;
    for (int p = threadIdx.x; p < L; p += blockDim.x) dst[p] = src[p];
    if (threadIdx.x == 0) {
      dst[L] = a.eot;
      a.fin_score[au * a.max_candidates + fin_slot[q]] = fin_sc[q];
      a.fin_len[au * a.max_candidates + fin_slot[q]] = L + 1;
    }
  }
  // next hypotheses: row s continues hypothesis c_src[keep[s]] with token c_tok[keep[s]] (through scratch rows: the
  // sources are rows of this audio, which this CTA is about to overwrite)
  for (int s2 = 0; s2 < n_keep; ++s2) {
    const int i = keep[s2];
    const int* src = a.tokens + static_cast<long long>(row0 + c_src[i]) * ld;
    int* dst = a.tokens_tmp + static_cast<long long>(row0 + s2) * ld;
    for (int p = threadIdx.x; p < L; p += blockDim.x) dst[p] = src[p];
    if (a.row_table != nullptr) {
      const int* ts = a.row_table + static_cast<long long>(row0 + c_src[i]) * a.table_ld;
      int* td = a.table_tmp + static_cast<long long>(row0 + s2) * a.table_ld;
      for (int p = threadIdx.x; p < L && p < a.table_ld; p += blockDim.x) td[p] = ts[p];
    }
  }
  __syncthreads();
  for (int s2 = 0; s2 < n_keep; ++s2) {
    const int i = keep[s2];
    const int* src = a.tokens_tmp + static_cast<long long>(row0 + s2) * ld;
    int* dst = a.tokens + static_cast<long long>(row0 + s2) * ld;
    for (int p = threadIdx.x; p < L; p += blockDim.x) dst[p] = src[p];
    if (a.row_table != nullptr) {
      const int* ts = a.table_tmp + static_cast<long long>(row0 + s2) * a.table_ld;
      int* td = a.row_table + static_cast<long long>(row0 + s2) * a.table_ld;
      for (int p = threadIdx.x; p < L && p < a.table_ld; p += blockDim.x) td[p] = ts[p];
    }
    if (threadIdx.x == 0) {
      dst[L] = c_tok[i];
      a.sum_logprobs_out[row0 + s2] = c_score[i];
    }
  }
  __syncthreads();
  for (int s2 = threadIdx.x; s2 < G; s2 += blockDim.x) {
    if (s2 < n_keep) a.sum_logprobs[row0 + s2] = a.sum_logprobs_out[row0 + s2];
    else a.sum_logprobs[row0 + s2] = -INFINITY;     // fewer than G live continuations: the row is dead
    a.hyp_id[row0 + s2] = a.R + row0 + s2;           // from now on every hypothesis is a distinct sequence
  }
}

int beam_step(const TopkArgs& tk, const BeamArgs& b, cudaStream_t stream) {
  WF_REQUIRE(b.G >= 1 && b.G <= 31 && b.R > 0 && b.R % b.G == 0, "beam step: beam size must be 1 .. 31 (got %d)", b.G);
  WF_REQUIRE(b.max_candidates >= 1 && b.max_candidates <= 32, "beam step: 1 .. 32 finished candidates per audio");
  WF_REQUIRE(b.tokens && b.tokens_tmp && b.state && b.sum_logprobs && b.sum_logprobs_out && b.hyp_id && b.vals && b.idx &&
                 b.fin_tokens && b.fin_score && b.fin_len && b.n_fin && (b.row_table == nullptr || b.table_tmp != nullptr),
             "beam step: null buffer");
  WF_REQUIRE(tk.state == b.state && tk.k == b.G + 1 && tk.out_vals == b.vals && tk.out_idx == b.idx,
             "beam step: the top-k pass must feed the update");
  int rc = topk_logprobs(tk, stream);
  if (rc) return rc;
  beam_update_kernel<<<b.R / b.G, 128, 0, stream>>>(b);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

// dst[r] = src[src_index[r]] for the first used_bytes of each row (beam reordering of the self-attention KV cache)
__global__ void __launch_bounds__(256)
kv_gather_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ idx, long long row_vec,
                 long long used_vec) {
  const int r = blockIdx.y;
  const uint4* s = src + static_cast<long long>(idx[r]) * row_vec;
  uint4* d = dst + static_cast<long long>(r) * row_vec;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < used_vec; i += stride)
    d[i] = s[i];
}
int kv_gather_rows(const void* src, void* dst, const int* src_index, int R, long long row_bytes, long long used_bytes,
                   cudaStream_t stream) {
  WF_REQUIRE(R > 0 && R <= 65535 && row_bytes % 16 == 0 && used_bytes % 16 == 0 && used_bytes <= row_bytes,
             "kv_gather: bad arguments");
  if (used_bytes == 0) return WF_OK;
  long long bx = (used_bytes / 16 + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(static_cast<unsigned>(bx), R);
  kv_gather_kernel<<<grid, 256, 0, stream>>>((const uint4*)src, (uint4*)dst, src_index, row_bytes / 16, used_bytes / 16);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

}  // namespace wf
