// Fused log-mel frontend:  PCM fp32 -> log-mel fp32   (reference whisper/audio.py:147-160)
//
//   reflect-pad 200 | 400-sample frames @ hop 160 | periodic Hann | rDFT (201 bins) | |X|^2 |
//   mel filterbank (80 / 128 rows, sparse) | log10(max(., 1e-10)) | max(., max - 8) | (. + 4) / 4
//
// Kernel 1 (logmel_fft_kernel) does everything up to log10 in one pass over the PCM: two real frames
// are packed into one complex 400-point FFT, computed in registers as 20 x 20 (each 20-point DFT is a
// twiddle-free 4 x 5 prime-factor transform), 20 threads per frame pair, one shared-memory transpose.
// It also reduces the running maximum (per clip and global) with warp shuffles + one atomic per block
// iteration.  Kernel 2 (logmel_finish_kernel) applies the max-8 clamp and the affine rescale in
// place; it is a pure streaming pass whose working set is normally still L2-resident.
// The intermediate 201 x 3000 complex spectrogram of the reference never touches HBM.
#include "common.cuh"
#include "kernels.h"
#include <vector>

namespace wf {

static constexpr int NFFT = 400;
static constexpr int HOP = 160;
static constexpr int NBINS = 201;
static constexpr int PAIRS = 16;             // frame pairs per block iteration
static constexpr int FRAMES_PER_GROUP = 2 * PAIRS;  // 32 frames -> 128-byte output rows
static constexpr int THREADS = PAIRS * 20;   // 320
static constexpr int SROW = 21;              // padded row of the 20x20 transpose (bank-conflict free)
static constexpr int MAX_W = 1024;

struct MelTables {
  float2 tw[NFFT];      // W400^j = (cos, -sin)(2 pi j / 400)
  float win[NFFT];      // periodic Hann
  short start[2][128];  // [set][mel]: first bin
  short count[2][128];  // number of bins
  short off[2][128];    // offset into w
  float w[2][MAX_W];
  int nnz[2];
};
__device__ MelTables g_tab;
static bool g_filters_set[PerDeviceOnce::MAX_DEV][2] = {};  // the __device__ table exists once per device
static bool g_consts_set = false;

int logmel_set_filters(int n_mels, const float* dense) {
  WF_REQUIRE(n_mels == 80 || n_mels == 128, "Unsupported n_mels: %d", n_mels);
  const int set = n_mels == 80 ? 0 : 1;
  static MelTables host;  // keeps both sets between calls
  if (!g_consts_set) {
    for (int j = 0; j < NFFT; ++j) {
      const double a = 2.0 * 3.14159265358979323846 * j / NFFT;
      host.tw[j] = make_float2((float)cos(a), (float)(-sin(a)));
      host.win[j] = (float)(0.5 - 0.5 * cos(a));
    }
    g_consts_set = true;
  }
  int off = 0;
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBINS; ++k)
      if (dense[m * NBINS + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
    if (lo < 0) { lo = 0; hi = -1; }
    host.start[set][m] = (short)lo;
    host.count[set][m] = (short)(hi - lo + 1);
    host.off[set][m] = (short)off;
    WF_REQUIRE(off + (hi - lo + 1) <= MAX_W, "mel filterbank too dense for the sparse table");
    for (int k = lo; k <= hi; ++k) host.w[set][off++] = dense[m * NBINS + k];
  }
  host.nnz[set] = off;
  // `host` accumulates both sets; the upload replaces the current device's whole table.  A set uploaded earlier to
  // ANOTHER device only is not valid here until it has been set on this device too.
  WF_CHECK_CUDA(cudaMemcpyToSymbol(g_tab, &host, sizeof(MelTables)));
  g_filters_set[current_device_slot()][set] = true;
  return WF_OK;
}

// ---- complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }

__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i t3
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i t3
}
__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  const float2 s14 = cadd(a1, a4), d14 = csub(a1, a4), s23 = cadd(a2, a3), d23 = csub(a2, a3);
  const float2 r1 = make_float2(a0.x + c1 * s14.x + c2 * s23.x, a0.y + c1 * s14.y + c2 * s23.y);
  const float2 r2 = make_float2(a0.x + c2 * s14.x + c1 * s23.x, a0.y + c2 * s14.y + c1 * s23.y);
  const float2 i1 = make_float2(s1 * d14.x + s2 * d23.x, s1 * d14.y + s2 * d23.y);
  const float2 i2 = make_float2(s2 * d14.x - s1 * d23.x, s2 * d14.y - s1 * d23.y);
  a0 = cadd(a0, cadd(s14, s23));
  a1 = make_float2(r1.x + i1.y, r1.y - i1.x);  // r1 - i i1
  a4 = make_float2(r1.x - i1.y, r1.y + i1.x);  // r1 + i i1
  a2 = make_float2(r2.x + i2.y, r2.y - i2.x);
  a3 = make_float2(r2.x - i2.y, r2.y + i2.x);
}
// 20-point forward DFT, prime-factor (Good-Thomas) 4 x 5: input index n = (5 n1 + 4 n2) mod 20,
// output index k = (5 k1 + 16 k2) mod 20; no internal twiddles.  o[] receives natural order.
__device__ __forceinline__ void dft20(float2 (&v)[20], float2 (&o)[20]) {
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2)
    dft4(v[(4 * n2) % 20], v[(5 + 4 * n2) % 20], v[(10 + 4 * n2) % 20], v[(15 + 4 * n2) % 20]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft5(v[(5 * k1) % 20], v[(5 * k1 + 4) % 20], v[(5 * k1 + 8) % 20], v[(5 * k1 + 12) % 20], v[(5 * k1 + 16) % 20]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) o[(5 * k1 + 16 * k2) % 20] = v[(5 * k1 + 4 * k2) % 20];
}

__device__ __forceinline__ int float_key(float f) {  // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ float load_reflect(const float* __restrict__ x, int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return __ldg(x + i);
}

template <int NMELS>
struct MelSmem {
  float2 tw[NFFT];
  float win[NFFT];
  float w[MAX_W];
  short start[NMELS], count[NMELS], off[NMELS];
  float2 S[PAIRS][20 * SROW];   // transpose buffer, later Z[400]
  float P[PAIRS][2][NBINS + 3];
  float tile[NMELS][FRAMES_PER_GROUP + 1];
  float red[THREADS / 32];
};

template <int NMELS>
__global__ void __launch_bounds__(THREADS, 2)
logmel_fft_kernel(const float* __restrict__ pcm, long long clip_stride, int n_samples, int n_frames, int n_clips,
                  int groups_per_clip, float* __restrict__ out, int* __restrict__ max_keys) {
  extern __shared__ uint8_t smem_raw[];
  MelSmem<NMELS>& sm = *reinterpret_cast<MelSmem<NMELS>*>(smem_raw);
  constexpr int SET = NMELS == 80 ? 0 : 1;
  const int tid = threadIdx.x;
  for (int i = tid; i < NFFT; i += THREADS) { sm.tw[i] = g_tab.tw[i]; sm.win[i] = g_tab.win[i]; }
  for (int i = tid; i < g_tab.nnz[SET]; i += THREADS) sm.w[i] = g_tab.w[SET][i];
  for (int i = tid; i < NMELS; i += THREADS) {
    sm.start[i] = g_tab.start[SET][i]; sm.count[i] = g_tab.count[SET][i]; sm.off[i] = g_tab.off[SET][i];
  }
  __syncthreads();

  const int pair = tid / 20, t = tid % 20;
  const int warp = tid >> 5, lane = tid & 31;
  float block_max = -3.0e38f;
  const long long total_groups = static_cast<long long>(n_clips) * groups_per_clip;

  for (long long g = blockIdx.x; g < total_groups; g += gridDim.x) {
    const int clip = static_cast<int>(g / groups_per_clip);
    const int f0 = static_cast<int>(g % groups_per_clip) * FRAMES_PER_GROUP;
    const float* x = pcm + clip * clip_stride;
    const int fa = f0 + 2 * pair, fb = fa + 1;

    // ---- load 2 windowed frames as one complex sequence z[n] = xa[n] + i xb[n]; thread t owns n = 20 n1 + t
    float2 v[20], o[20];
    {
      const int base_a = fa * HOP - NFFT / 2, base_b = base_a + HOP;
      const bool va = fa < n_frames, vb = fb < n_frames;
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const int j = 20 * n1 + t;
        const float w = sm.win[j];
        v[n1].x = va ? load_reflect(x, base_a + j, n_samples) * w : 0.f;
        v[n1].y = vb ? load_reflect(x, base_b + j, n_samples) * w : 0.f;
      }
    }
    // ---- stage 1: 20-point DFT over n1 (for fixed n2 = t), twiddle by W400^(n2 k1), transpose through smem
    dft20(v, o);
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) sm.S[pair][k1 * SROW + t] = cmul(o[k1], sm.tw[t * k1]);
    __syncthreads();
    // ---- stage 2: thread t = k1 gathers its row and transforms over n2 -> Z[k1 + 20 k2]
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = sm.S[pair][t * SROW + n2];
    __syncthreads();
    dft20(v, o);
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) sm.S[pair][t + 20 * k2] = o[k2];
    __syncthreads();
    // ---- split the two real spectra and take |X|^2 for bins 0..200
    for (int k = t; k < NBINS; k += 20) {
      const float2 z = sm.S[pair][k];
      const float2 w = sm.S[pair][(NFFT - k) % NFFT];
      const float ar = z.x + w.x, ai = z.y - w.y;  // 2 Xa
      const float br = z.y + w.y, bi = z.x - w.x;  // 2 Xb (up to sign)
      sm.P[pair][0][k] = 0.25f * (ar * ar + ai * ai);
      sm.P[pair][1][k] = 0.25f * (br * br + bi * bi);
    }
    __syncthreads();
    // ---- mel projection + log10 into the [NMELS][32] output tile
    for (int idx = t; idx < 2 * NMELS; idx += 20) {
      const int f = idx / NMELS, m = idx % NMELS;
      const float* p = &sm.P[pair][f][sm.start[m]];
      const float* w = &sm.w[sm.off[m]];
      float acc = 0.f;
      for (int j = 0; j < sm.count[m]; ++j) acc = fmaf(w[j], p[j], acc);
      sm.tile[m][2 * pair + f] = log10f(fmaxf(acc, 1e-10f));
    }
    __syncthreads();
    // ---- coalesced store (one 128-byte row segment per warp instruction) + running max
    float mx = -3.0e38f;
    {
      const int f = f0 + lane;
      if (f < n_frames) {
        float* dst = out + (static_cast<long long>(clip) * NMELS) * n_frames + f;
        for (int m = warp; m < NMELS; m += THREADS / 32) {
          const float val = sm.tile[m][lane];
          dst[static_cast<long long>(m) * n_frames] = val;
          mx = fmaxf(mx, val);
        }
      }
    }
    mx = warp_max(mx);
    if (lane == 0) sm.red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      float m2 = sm.red[0];
#pragma unroll
      for (int i = 1; i < THREADS / 32; ++i) m2 = fmaxf(m2, sm.red[i]);
      atomicMax(max_keys + clip, float_key(m2));
      block_max = fmaxf(block_max, m2);
    }
    // (sm.red / sm.tile are rewritten only after the next iteration's barriers)
  }
  if (tid == 0 && block_max > -3.0e38f) atomicMax(max_keys + n_clips, float_key(block_max));
}

// out = (max(out, mx - 8) + 4) / 4 with mx = per-clip max (mode 1) or whole-tensor max (mode 0, the
// reference's semantics for batched input: audio.py:159).
__global__ void __launch_bounds__(256)
logmel_finish_kernel(float* __restrict__ out, long long per_clip, int n_clips, const int* __restrict__ max_keys,
                     int mode) {
  const int clip = blockIdx.y;
  const float mx = key_float(mode == 1 ? max_keys[clip] : max_keys[n_clips]);
  const float floor_v = mx - 8.0f;
  float* p = out + clip * per_clip;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((per_clip & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const long long n4 = per_clip >> 2;
    for (; i < n4; i += stride) {
      float4 v = p4[i];
      v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
      v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
      v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
      v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
      p4[i] = v;
    }
  } else {
    for (; i < per_clip; i += stride) p[i] = (fmaxf(p[i], floor_v) + 4.0f) * 0.25f;
  }
}

long long logmel_workspace_bytes(int n_clips) { return static_cast<long long>(n_clips + 1) * sizeof(int); }

template <int NMELS>
static int launch_logmel(const float* pcm, int n_clips, int n_samples, long long clip_stride, int mode, float* out,
                         int* keys, cudaStream_t stream) {
  const int n_frames = n_samples / HOP;
  const int groups = (n_frames + FRAMES_PER_GROUP - 1) / FRAMES_PER_GROUP;
  const long long total = static_cast<long long>(n_clips) * groups;
  const int smem = static_cast<int>(sizeof(MelSmem<NMELS>));
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(logmel_fft_kernel<NMELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  WF_CHECK_CUDA(cudaMemsetAsync(keys, 0x80, (n_clips + 1) * sizeof(int), stream));
  const long long max_grid = 2LL * num_sms();
  const int grid = static_cast<int>(total < max_grid ? total : max_grid);
  logmel_fft_kernel<NMELS><<<grid, THREADS, smem, stream>>>(pcm, clip_stride, n_samples, n_frames, n_clips, groups,
                                                           out, keys);
  WF_CHECK_LAUNCH();
  const long long per_clip = static_cast<long long>(NMELS) * n_frames;
  long long bx = (per_clip / 4 + 255) / 256;
  const long long want = (4LL * num_sms() + n_clips - 1) / n_clips;
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  dim3 fgrid(static_cast<unsigned>(bx), static_cast<unsigned>(n_clips));
  logmel_finish_kernel<<<fgrid, 256, 0, stream>>>(out, per_clip, n_clips, keys, mode);
  WF_CHECK_LAUNCH();
  return WF_OK;
}

int logmel_f32(const float* pcm, int n_clips, int n_samples, long long clip_stride, int n_mels, int mode, float* out,
               void* workspace, cudaStream_t stream) {
  WF_REQUIRE(n_mels == 80 || n_mels == 128, "Unsupported n_mels: %d", n_mels);
  WF_REQUIRE(g_filters_set[current_device_slot()][n_mels == 80 ? 0 : 1],
             "wf_logmel_set_filters(%d) has not been called on this device", n_mels);
  WF_REQUIRE(n_clips > 0 && n_clips <= 65535, "logmel: n_clips=%d out of range [1, 65535]", n_clips);
  WF_REQUIRE(n_samples > NFFT / 2, "logmel: reflect padding needs more than %d samples (got %d)", NFFT / 2, n_samples);
  WF_REQUIRE(n_samples / HOP > 0, "logmel: no complete frame in %d samples", n_samples);
  WF_REQUIRE(mode == 0 || mode == 1, "logmel: mode must be 0 (global max) or 1 (per-clip max)");
  WF_REQUIRE(workspace != nullptr, "logmel: workspace is null");
  int* keys = reinterpret_cast<int*>(workspace);
  return n_mels == 80 ? launch_logmel<80>(pcm, n_clips, n_samples, clip_stride, mode, out, keys, stream)
                      : launch_logmel<128>(pcm, n_clips, n_samples, clip_stride, mode, out, keys, stream);
}

}  // namespace wf
