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
//
// logmel_fft2_kernel (round 2, the default) is the same arithmetic laid out the other way round: a LANE is a frame
// pair and a WARP is a DFT row.  A CTA of 20 warps takes 64 consecutive frames (lane p: frames f0+p and f0+32+p as
// the real and imaginary part of one complex sequence); warp t runs the 20-point column transform n = 20 n1 + t of
// 32 pairs at once, then warp k1 the row transform.  Everything indexed by t, k1, n2 or the mel row is therefore
// warp-uniform: window, twiddles and filter weights are broadcast shared-memory loads at immediate offsets, the
// transposes are conflict-free [index][lane] arrays, the output row of a mel bin is one coalesced 128-byte store
// straight from registers, and no per-element integer arithmetic is left (the first kernel spent 44 % of its issue
// slots on it: 20-thread groups straddling warps, reflect arithmetic on all 800 loads of a pair, a scalar sparse
// filterbank walk).  The PCM of the next group is staged with cp.async (4-byte, into hop rows padded to 161 words so
// that a warp's 32 frames hit 32 banks) while the current one is transformed; only the two real spectra's upper
// halves cross warps (row k1 needs row 20-k1), the power spectrum goes through shared memory once for the filterbank.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <utility>
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
  // ---- second layout (logmel_fft2_kernel)
  float2 twt[20][20];          // twt[t][k1] = W400^(t k1)
  float wint[20][20];          // wint[t][n1] = hann[20 n1 + t]
  float w4[2][MAX_W];          // filter rows padded to multiples of 4 weights, scaled by 1/4 (|2X|^2 -> |X|^2)
  short cnt4[2][128];          // quads per row
  short off4[2][128];          // offset of a row in w4 (multiple of 4)
  short assign[2][20][8];      // mel rows of warp w (snake deal by length, -1 terminated)
  int nnz4[2];
  // ---- third layout: the rows of the two half-warps of a warp are padded to a common number of quads
  float w3[2][MAX_W];
  short off3[2][128];
  short nq3[2][128];
  int nnz3[2];
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
    for (int t = 0; t < 20; ++t)
      for (int i = 0; i < 20; ++i) {
        host.twt[t][i] = host.tw[(t * i) % NFFT];
        host.wint[t][i] = host.win[20 * i + t];
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
  {  // padded rows + a balanced deal of the rows to the 20 warps (longest first, snake order)
    int o4 = 0;
    std::vector<int> order(n_mels);
    for (int m = 0; m < n_mels; ++m) {
      const int cnt = host.count[set][m], q = (cnt + 3) / 4;
      WF_REQUIRE(o4 + 4 * q <= MAX_W, "mel filterbank too dense for the padded table");
      host.off4[set][m] = (short)o4;
      host.cnt4[set][m] = (short)q;
      for (int j = 0; j < 4 * q; ++j)
        host.w4[set][o4 + j] = j < cnt ? 0.25f * host.w[set][host.off[set][m] + j] : 0.f;
      o4 += 4 * q;
      order[m] = m;
    }
    host.nnz4[set] = o4;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return host.cnt4[set][a] > host.cnt4[set][b]; });
    WF_REQUIRE(n_mels <= 20 * 8, "too many mel rows for the warp assignment table");
    for (int w = 0; w < 20; ++w)
      for (int s2 = 0; s2 < 8; ++s2) host.assign[set][w][s2] = -1;
    for (int i = 0; i < n_mels; ++i) {
      const int round = i / 20, pos = i % 20;
      host.assign[set][(round & 1) ? 19 - pos : pos][round] = (short)order[i];
    }
    // third layout: virtual warps 2 w and 2 w + 1 walk their slot-s rows in lock step
    int o3 = 0;
    for (int vw = 0; vw < 20; vw += 2)
      for (int s2 = 0; s2 < 8; ++s2) {
        const int m0 = host.assign[set][vw][s2], m1 = host.assign[set][vw + 1][s2];
        const int q = std::max(m0 >= 0 ? (int)host.cnt4[set][m0] : 0, m1 >= 0 ? (int)host.cnt4[set][m1] : 0);
        for (int m : {m0, m1}) {
          if (m < 0) continue;
          WF_REQUIRE(o3 + 4 * q <= MAX_W, "mel filterbank too dense for the paired table");
          WF_REQUIRE(host.start[set][m] + 4 * q <= 236, "mel filterbank rows too uneven for the paired table");
          host.off3[set][m] = (short)o3;
          host.nq3[set][m] = (short)q;
          const int own = 4 * host.cnt4[set][m];
          for (int j = 0; j < 4 * q; ++j) host.w3[set][o3 + j] = j < own ? host.w4[set][host.off4[set][m] + j] : 0.f;
          o3 += 4 * q;
        }
      }
    host.nnz3[set] = o3;
  }
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

// ------------------------------------------------------------------------------------------------------------------
// Second layout: lane = frame pair, warp = DFT row (see the header).  Complex values travel as packed f32x2 register
// pairs: Blackwell's FADD2 / FMUL2 / FFMA2 do both components in one issue slot.
static constexpr int NW2 = 20;                   // warps per CTA = rows / columns of the 20 x 20 transform
static constexpr int TH2 = NW2 * 32;             // 640 threads
static constexpr int GF2 = 64;                   // frames per group
static constexpr int PROW = 161;                 // padded hop row of the staged PCM (161 = 1 mod 32)
static constexpr int PCM_ROWS = 66;
static constexpr int GSAMP = (GF2 - 1) * HOP + NFFT;   // 10480 samples feed 64 frames
static constexpr int PK = 204;                   // power-spectrum rows (201 bins + quad padding)
static constexpr int SLOTS = 8;                  // mel rows per warp, at most

typedef unsigned long long c2;                   // (re, im) or (frame a, frame b) as one f32x2 operand
__device__ __forceinline__ c2 pk(float x, float y) { c2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void upk(c2 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ c2 add2(c2 a, c2 b) { c2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 sub2(c2 a, c2 b) { c2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 mul2(c2 a, c2 b) { c2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c) {
  c2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
// -i (a - b) = (a.y - b.y, b.x - a.x): the one place where the components cross
__device__ __forceinline__ c2 rotsub(c2 a, c2 b) {
  float ax, ay, bx, by; upk(a, ax, ay); upk(b, bx, by);
  return pk(ay - by, bx - ax);
}
struct Dft5Consts { c2 c1, c2_, s1, s2, ns1; };
__device__ __forceinline__ void dft4p(c2& a0, c2& a1, c2& a2, c2& a3) {
  const c2 t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3), t3r = rotsub(a1, a3);
  a0 = add2(t0, t2);
  a2 = sub2(t0, t2);
  a1 = add2(t1, t3r);   // t1 - i t3
  a3 = sub2(t1, t3r);   // t1 + i t3
}
__device__ __forceinline__ void dft5p(c2& a0, c2& a1, c2& a2, c2& a3, c2& a4, const Dft5Consts& k) {
  const c2 s14 = add2(a1, a4), s23 = add2(a2, a3), d14r = rotsub(a1, a4), d23r = rotsub(a2, a3);
  const c2 r1 = fma2(s23, k.c2_, fma2(s14, k.c1, a0));
  const c2 r2 = fma2(s23, k.c1, fma2(s14, k.c2_, a0));
  const c2 i1r = fma2(d23r, k.s2, mul2(d14r, k.s1));    // -i (s1 d14 + s2 d23)
  const c2 i2r = fma2(d23r, k.ns1, mul2(d14r, k.s2));   // -i (s2 d14 - s1 d23)
  a0 = add2(a0, add2(s14, s23));
  a1 = add2(r1, i1r);
  a4 = sub2(r1, i1r);
  a2 = add2(r2, i2r);
  a3 = sub2(r2, i2r);
}
// the 20-point prime-factor transform of dft20() on packed values
__device__ __forceinline__ void dft20p(c2 (&v)[20], c2 (&o)[20], const Dft5Consts& k) {
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2)
    dft4p(v[(4 * n2) % 20], v[(5 + 4 * n2) % 20], v[(10 + 4 * n2) % 20], v[(15 + 4 * n2) % 20]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft5p(v[(5 * k1) % 20], v[(5 * k1 + 4) % 20], v[(5 * k1 + 8) % 20], v[(5 * k1 + 12) % 20], v[(5 * k1 + 16) % 20], k);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) o[(5 * k1 + 16 * k2) % 20] = v[(5 * k1 + 4 * k2) % 20];
}

template <int NMELS>
struct MelSmem2 {
  c2 S[NFFT * 32];              // [k1][n2][lane] between the stages; its first half is reused as X[row][kk][lane]
  c2 P[PK * 32];                // [bin][lane]: (|2 X_a|^2, |2 X_b|^2)
  float pcm[PCM_ROWS * PROW];
  float2 twt[NFFT];             // [t][k1]
  c2 win2[NFFT];                // [t][n1]: (h, h)
  __align__(16) c2 w4[MAX_W];   // (w, w), rows padded to quads
  int4 slot[NW2][SLOTS];        // {byte offset of the first bin in P, byte offset of the row in w4, quads, mel row or -1}
};

template <int Q>
__device__ __forceinline__ void cp_async4_q(uint32_t dst, const float* src) {   // element tid + 640 Q of a group
  asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 4;" ::"r"(dst), "l"(src), "n"(Q * 4 * PROW * 4),
               "n"(Q * TH2 * 4) : "memory");
}
template <int... Q>
__device__ __forceinline__ void cp_async4_seq(uint32_t dst, const float* src, std::integer_sequence<int, Q...>) {
  (cp_async4_q<Q>(dst, src), ...);
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float log10_fast(float x) {   // x >= 1e-10: lg2.approx is within 2^-22 of log2
  float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return 0.30102999566398120f * r;
}

template <int NMELS>
__global__ void __launch_bounds__(TH2, 1)
logmel_fft2_kernel(const float* __restrict__ pcm, long long clip_stride, int n_samples, int n_frames, int n_clips,
                   int groups_per_clip, float* __restrict__ out, int* __restrict__ max_keys) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem2<NMELS>& sm = *reinterpret_cast<MelSmem2<NMELS>*>(smem_raw);
  constexpr int SET = NMELS == 80 ? 0 : 1;
  const int tid = threadIdx.x;
  const int w = tid >> 5, p = tid & 31;
  for (int i = tid; i < NFFT; i += TH2) {
    sm.twt[i] = g_tab.twt[i / 20][i % 20];
    const float h = g_tab.wint[i / 20][i % 20];
    sm.win2[i] = pk(h, h);
  }
  for (int i = tid; i < g_tab.nnz4[SET]; i += TH2) { const float c = g_tab.w4[SET][i]; sm.w4[i] = pk(c, c); }
  for (int i = tid; i < NW2 * SLOTS; i += TH2) {
    const int m = g_tab.assign[SET][i / SLOTS][i % SLOTS];
    int4 d = make_int4(0, 0, 0, -1);
    if (m >= 0) d = make_int4(g_tab.start[SET][m] * 32 * 8, g_tab.off4[SET][m] * 8, g_tab.cnt4[SET][m], m);
    sm.slot[i / SLOTS][i % SLOTS] = d;
  }
  for (int i = tid; i < 3 * 32; i += TH2) sm.P[NBINS * 32 + i] = 0ull;   // the quad padding reads bins 201..203

  const long long total_groups = static_cast<long long>(n_clips) * groups_per_clip;
  const long long g_begin = total_groups * blockIdx.x / gridDim.x;
  const long long g_end = total_groups * (blockIdx.x + 1) / gridDim.x;

  // stage the 10480 samples of group g: element i = tid + 640 q sits in hop row tid / 160 + 4 q at column tid % 160
  const int col0 = tid % HOP, row0 = tid / HOP;
  const uint32_t dst0 = smem_u32(sm.pcm + row0 * PROW + col0);
  auto stage = [&](long long g) {
    const int clip = static_cast<int>(g / groups_per_clip);
    const int f0 = static_cast<int>(g % groups_per_clip) * GF2;
    const float* x = pcm + clip * clip_stride;
    const int gbase = f0 * HOP - NFFT / 2;
    if (gbase >= 0 && gbase + GSAMP <= n_samples) {
      const float* src = x + gbase + tid;
      cp_async4_seq(dst0, src, std::make_integer_sequence<int, 16>{});
      if (tid < GSAMP - 16 * TH2) cp_async4_q<16>(dst0, src);
    } else {  // first / last group of a clip: reflect at both ends (frames past n_frames read clamped garbage)
#pragma unroll 1
      for (int q = 0; q < 17; ++q) {
        const int i = tid + q * TH2;
        if (i < GSAMP) {
          int idx = gbase + i;
          if (idx < 0) idx = -idx;
          if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
          idx = max(0, min(idx, n_samples - 1));
          cp_async4(dst0 + q * (4 * PROW * 4), x + idx);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  Dft5Consts kc;
  kc.c1 = pk(0.30901699437494745f, 0.30901699437494745f);
  kc.c2_ = pk(-0.80901699437494745f, -0.80901699437494745f);
  kc.s1 = pk(0.95105651629515353f, 0.95105651629515353f);
  kc.s2 = pk(0.58778525229247314f, 0.58778525229247314f);
  kc.ns1 = pk(-0.95105651629515353f, -0.95105651629515353f);

  float run_max = -3.0e38f, all_max = -3.0e38f;
  int cur_clip = -1;
  auto flush = [&]() {
    const float m = warp_max(run_max);
    if (p == 0 && cur_clip >= 0 && m > -3.0e38f) atomicMax(max_keys + cur_clip, float_key(m));
    all_max = fmaxf(all_max, m);
  };

  __syncthreads();
  if (g_begin < g_end) stage(g_begin);

  for (long long g = g_begin; g < g_end; ++g) {
    const int clip = static_cast<int>(g / groups_per_clip);
    const int f0 = static_cast<int>(g % groups_per_clip) * GF2;
    if (clip != cur_clip) { flush(); cur_clip = clip; run_max = -3.0e38f; }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                                       // (A) PCM of this group is in place

    c2 v[20], o[20];
    // ---- stage 1: warp w = column t; z[n] = xa[n] + i xb[n], n = 20 n1 + t, windowed
    {
      const float* ps = sm.pcm + p * PROW + w;
      const c2* wn = sm.win2 + w * 20;
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const int off = (n1 / 8) * PROW + 20 * (n1 % 8);
        v[n1] = mul2(pk(ps[off], ps[off + 32 * PROW]), wn[n1]);
      }
    }
    dft20p(v, o, kc);
    {
      const float2* tw = sm.twt + w * 20;
      c2* dst = sm.S + w * 32 + p;
#pragma unroll
      for (int k1 = 0; k1 < 20; ++k1) {
        float x, y; upk(o[k1], x, y);
        const float2 t = tw[k1];
        dst[k1 * 640] = pk(x * t.x - y * t.y, x * t.y + y * t.x);
      }
    }
    __syncthreads();                                                       // (B)
    if (g + 1 < g_end) stage(g + 1);                                        // every read of sm.pcm is done
    // ---- stage 2: warp w = row k1 gathers its 20 columns
    {
      const c2* src = sm.S + w * 640 + p;
#pragma unroll
      for (int n2 = 0; n2 < 20; ++n2) v[n2] = src[n2 * 32];
    }
    __syncthreads();                                                       // (C) S may be overwritten
    dft20p(v, o, kc);                                                      // o[k2] = Z[w + 20 k2]
    {
      c2* X = sm.S + w * 320 + p;
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) X[kk * 32] = o[10 + kk];
    }
    __syncthreads();                                                       // (D)
    // ---- the two real spectra: 2 X_a[k] = Z[k] + conj Z[400-k], 2 X_b[k] = -i (Z[k] - conj Z[400-k]); |2X|^2 to P
    {
      c2* pp = sm.P + w * 32 + p;
      auto power = [&](c2 z, c2 c) {   // with c = Z[400-k]: (ar, br) = z + c, (bi, ai) = z - c
        float ar, br, bi2, ai2;
        upk(add2(z, c), ar, br);
        const c2 d = sub2(z, c);
        upk(mul2(d, d), bi2, ai2);
        return pk(fmaf(ar, ar, ai2), fmaf(br, br, bi2));
      };
      if (w == 0) {   // bins 20 k2, k2 = 0..10; the partner Z[400 - 20 k2] is in this row too
        const c2* X = sm.S + p;
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) pp[k2 * 640] = power(o[k2], k2 == 0 ? o[0] : X[(10 - k2) * 32]);
      } else {        // bins w + 20 k2, k2 = 0..9; partner = row 20 - w, element 19 - k2
        const c2* X = sm.S + (20 - w) * 320 + p;
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) pp[k2 * 640] = power(o[k2], X[(9 - k2) * 32]);
      }
    }
    __syncthreads();                                                       // (E)
    // ---- mel rows of this warp: filterbank, log10, store, running maximum
    {
      const int fa = f0 + p;
      const bool oka = fa < n_frames, okb = fa + 32 < n_frames;
      float* dst = out + (static_cast<long long>(clip) * NMELS) * n_frames + fa;
      const uint32_t pbase = smem_u32(sm.P + p), wbase = smem_u32(sm.w4);
#pragma unroll 1
      for (int s = 0; s < SLOTS; ++s) {
        const int4 d = sm.slot[w][s];
        if (d.w < 0) break;
        uint32_t qa = pbase + d.x, wq = wbase + d.y;
        c2 acc0 = 0ull, acc1 = 0ull;
#pragma unroll 1
        for (int j = 0; j < d.z; ++j) {
          c2 c0, c1, c2v, c3, q0, q1, q2, q3;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "r"(wq));
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2 + 16];" : "=l"(c2v), "=l"(c3) : "r"(wq));
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(q0) : "r"(qa));
          asm volatile("ld.shared.b64 %0, [%1 + 256];" : "=l"(q1) : "r"(qa));
          asm volatile("ld.shared.b64 %0, [%1 + 512];" : "=l"(q2) : "r"(qa));
          asm volatile("ld.shared.b64 %0, [%1 + 768];" : "=l"(q3) : "r"(qa));
          acc0 = fma2(c0, q0, acc0);
          acc1 = fma2(c1, q1, acc1);
          acc0 = fma2(c2v, q2, acc0);
          acc1 = fma2(c3, q3, acc1);
          qa += 1024; wq += 32;
        }
        float ma, mb;
        upk(add2(acc0, acc1), ma, mb);
        const float va = log10_fast(fmaxf(ma, 1e-10f)), vb = log10_fast(fmaxf(mb, 1e-10f));
        float* o2 = dst + static_cast<long long>(d.w) * n_frames;
        if (oka) { o2[0] = va; run_max = fmaxf(run_max, va); }
        if (okb) { o2[32] = vb; run_max = fmaxf(run_max, vb); }
      }
    }
  }
  flush();
  if (p == 0 && all_max > -3.0e38f) atomicMax(max_keys + n_clips, float_key(all_max));
}

// ------------------------------------------------------------------------------------------------------------------
// Third layout = the second one at half width, two CTAs per SM.  The 20-warp kernel runs its phases in lock step (one
// CTA per SM, five barriers per group) and is bound by the shared-memory data pipe (ncu: LSU wavefronts 67 % of peak,
// issue slots 45 % busy, MIO throttle + short scoreboard the top stalls).  Here a CTA is 10 warps and takes 32 frames
// (16 pairs); a warp's two half-warps are two DFT columns (t = 2 w + h) and then the two ROWS r and 20 - r of the
// second stage (warp 0: rows 0 and 10, their own partners), so the conjugate-partner exchange of the real-spectrum
// split is a lane-xor-16 shuffle instead of a trip through shared memory.  The power spectrum has its own buffer:
// two barriers per group, 110 KB per CTA, two CTAs with independent barriers per SM.
// Measured and rejected: applying the clamp + rescale inside this kernel (device-wide group queue, groups counted per
// clip, the CTA that completes a clip rewrites its 1 MB from L2): one CTA needs ~70 us per clip whatever the number of
// loads in flight, and the completions pile up at the end of the grid - 338 us against 226 us at 128 clips, 1866 against
// 1745 us at 1024.  The second pass stays a kernel of its own (HBM-bound: it reads and writes the output once more).
static constexpr int NW3 = 10;
static constexpr int TH3 = NW3 * 32;             // 320 threads
static constexpr int GF3 = 32;                   // frames per group: lane (h, p) holds frames f0 + p and f0 + 16 + p
static constexpr int PROW3 = 162;                // hop row padded to 2 mod 32: bank = 2 p + t + const, all 32 distinct
static constexpr int PCM_ROWS3 = 34;
static constexpr int GSAMP3 = (GF3 - 1) * HOP + NFFT;   // 5360 samples feed 32 frames
static constexpr int PK3 = 236;                  // power-spectrum rows incl. the padding the paired quad walk may read

template <int NMELS>
struct MelSmem3 {
  c2 S[NFFT * 16];              // [k1][n2][16] between the stages
  c2 P[PK3 * 16];               // [bin][16]: (|2 X_a|^2, |2 X_b|^2); rows from 201 on stay zero (quad padding)
  float pcm[PCM_ROWS3 * PROW3];
  __align__(16) float2 twt[NFFT];   // [t][k1]
  __align__(16) float win[NFFT];    // [t][n1]
  __align__(16) float w4[MAX_W];    // rows padded to quads
  int4 slot[NW3][SLOTS][2];     // per half-warp: {byte offset of the first bin in P, byte offset in w4, quads, mel row}
};

template <int Q>
__device__ __forceinline__ void cp_async4_q3(uint32_t dst, const float* src) {   // element tid + 320 Q of a group
  asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 4;" ::"r"(dst), "l"(src), "n"(Q * 2 * PROW3 * 4),
               "n"(Q * TH3 * 4) : "memory");
}
template <int... Q>
__device__ __forceinline__ void cp_async4_seq3(uint32_t dst, const float* src, std::integer_sequence<int, Q...>) {
  (cp_async4_q3<Q>(dst, src), ...);
}
__device__ __forceinline__ c2 shfl16(c2 v) { return __shfl_xor_sync(0xffffffffu, v, 16); }
// (|2 X_a|^2, |2 X_b|^2) from z = Z[k] and c = Z[400 - k]: (ar, br) = z + c, (bi, ai) = z - c
__device__ __forceinline__ c2 power2(c2 z, c2 c) {
  float ar, br, bi2, ai2;
  upk(add2(z, c), ar, br);
  const c2 d = sub2(z, c);
  upk(mul2(d, d), bi2, ai2);
  return pk(fmaf(ar, ar, ai2), fmaf(br, br, bi2));
}

template <int NMELS>
__global__ void __launch_bounds__(TH3, 2)
logmel_fft3_kernel(const float* __restrict__ pcm, long long clip_stride, int n_samples, int n_frames, int n_clips,
                   int groups_per_clip, float* __restrict__ out, int* __restrict__ max_keys) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem3<NMELS>& sm = *reinterpret_cast<MelSmem3<NMELS>*>(smem_raw);
  constexpr int SET = NMELS == 80 ? 0 : 1;
  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31, p = lane & 15, h = lane >> 4;
  const int t = 2 * w + h;                                       // column of the first stage
  const int r = w == 0 ? 10 * h : (h ? 20 - w : w);             // row of the second stage; the other half has 20 - r
  for (int i = tid; i < NFFT; i += TH3) {
    sm.twt[i] = g_tab.twt[i / 20][i % 20];
    sm.win[i] = g_tab.wint[i / 20][i % 20];
  }
  for (int i = tid; i < g_tab.nnz3[SET]; i += TH3) sm.w4[i] = g_tab.w3[SET][i];
  for (int i = tid; i < 20 * SLOTS; i += TH3) {   // virtual warp 2 w + h of the 20-way deal
    const int vw = i / SLOTS, sl = i % SLOTS;
    const int m = g_tab.assign[SET][vw][sl];
    int4 d = make_int4(0, 0, 0, -1);
    if (m >= 0) d = make_int4(g_tab.start[SET][m] * 16 * 8, g_tab.off3[SET][m] * 4, g_tab.nq3[SET][m], m);
    sm.slot[vw >> 1][sl][vw & 1] = d;
  }
  for (int i = tid; i < (PK3 - NBINS) * 16; i += TH3) sm.P[NBINS * 16 + i] = 0ull;

  const long long total_groups = static_cast<long long>(n_clips) * groups_per_clip;
  const long long g_begin = total_groups * blockIdx.x / gridDim.x;
  const int n_groups = static_cast<int>(total_groups * (blockIdx.x + 1) / gridDim.x - g_begin);
  int clip = static_cast<int>(g_begin / groups_per_clip);     // of the group being transformed
  int gi = static_cast<int>(g_begin % groups_per_clip);

  // stage the 5360 samples of group g: element i = tid + 320 q sits in hop row tid / 160 + 2 q at column tid % 160
  const int col0 = tid % HOP, row0 = tid / HOP;
  const uint32_t dst0 = smem_u32(sm.pcm + row0 * PROW3 + col0);
  auto stage = [&](int sclip, int sgi) {
    const int f0 = sgi * GF3;
    const float* x = pcm + sclip * clip_stride;
    const int gbase = f0 * HOP - NFFT / 2;
    if (gbase >= 0 && gbase + GSAMP3 <= n_samples) {
      const float* src = x + gbase + tid;
      cp_async4_seq3(dst0, src, std::make_integer_sequence<int, 16>{});
      if (tid < GSAMP3 - 16 * TH3) cp_async4_q3<16>(dst0, src);
    } else {  // first / last group of a clip: reflect at both ends (frames past n_frames read clamped garbage)
#pragma unroll 1
      for (int q = 0; q < 17; ++q) {
        const int i = tid + q * TH3;
        if (i < GSAMP3) {
          int idx = gbase + i;
          if (idx < 0) idx = -idx;
          if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
          idx = max(0, min(idx, n_samples - 1));
          cp_async4(dst0 + q * (2 * PROW3 * 4), x + idx);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  Dft5Consts kc;
  kc.c1 = pk(0.30901699437494745f, 0.30901699437494745f);
  kc.c2_ = pk(-0.80901699437494745f, -0.80901699437494745f);
  kc.s1 = pk(0.95105651629515353f, 0.95105651629515353f);
  kc.s2 = pk(0.58778525229247314f, 0.58778525229247314f);
  kc.ns1 = pk(-0.95105651629515353f, -0.95105651629515353f);

  float run_max = -3.0e38f, all_max = -3.0e38f;
  int cur_clip = -1;
  auto flush = [&]() {
    const float m = warp_max(run_max);
    if (lane == 0 && cur_clip >= 0 && m > -3.0e38f) atomicMax(max_keys + cur_clip, float_key(m));
    all_max = fmaxf(all_max, m);
  };

  __syncthreads();
  if (n_groups > 0) stage(clip, gi);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();                                     // PCM of the first group is in place

  for (int it = 0; it < n_groups; ++it) {
    const int f0 = gi * GF3;
    if (clip != cur_clip) { flush(); cur_clip = clip; run_max = -3.0e38f; }
    // (no barrier here: barrier E of the previous group already orders everything the next stages touch - every
    // thread waited for its PCM copies before it, the S reads of the second stage and the P writes precede it, and
    // the next P writes come after barrier B, which no warp passes before all have left the filterbank)

    c2 v[20], o[20];
    // ---- stage 1: half-warp = column t; z[n] = xa[n] + i xb[n], n = 20 n1 + t, windowed
    {
      const float* ps = sm.pcm + p * PROW3 + t;
      const float4* wn = reinterpret_cast<const float4*>(sm.win + t * 20);
#pragma unroll
      for (int n4 = 0; n4 < 5; ++n4) {
        const float4 hq = wn[n4];
        const float hw[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n1 = 4 * n4 + e;
          const int off = (n1 / 8) * PROW3 + 20 * (n1 % 8);
          v[n1] = pk(ps[off] * hw[e], ps[off + 16 * PROW3] * hw[e]);
        }
      }
    }
    dft20p(v, o, kc);
    {
      const float4* tw = reinterpret_cast<const float4*>(sm.twt + t * 20);
      c2* dst = sm.S + t * 16 + p;
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        const float4 f = tw[k2];   // twiddles of k1 = 2 k2 and 2 k2 + 1
        float x, y;
        upk(o[2 * k2], x, y);
        dst[(2 * k2) * 320] = pk(x * f.x - y * f.y, x * f.y + y * f.x);
        upk(o[2 * k2 + 1], x, y);
        dst[(2 * k2 + 1) * 320] = pk(x * f.z - y * f.w, x * f.w + y * f.z);
      }
    }
    __syncthreads();                                   // (B)
    const int out_clip = clip;
    if (++gi == groups_per_clip) { gi = 0; ++clip; }   // the next group
    if (it + 1 < n_groups) stage(clip, gi);             // every read of sm.pcm is done
    // ---- stage 2: half-warp = row r gathers its 20 columns
    {
      const c2* src = sm.S + r * 320 + p;
#pragma unroll
      for (int n2 = 0; n2 < 20; ++n2) v[n2] = src[n2 * 16];
    }
    dft20p(v, o, kc);                                  // o[k2] = Z[r + 20 k2]
    // ---- the two real spectra: 2 X_a[k] = Z[k] + conj Z[400-k], 2 X_b[k] = -i (Z[k] - conj Z[400-k]); |2X|^2 to P.
    // Row r holds bins r + 20 k2 (k2 < 10); the partner Z[400 - k] is element 19 - k2 of row 20 - r = the other
    // half-warp.  Rows 0 and 10 (warp 0) are their own partners: element (20 - k2) % 20 in row 0, 19 - k2 in row 10,
    // and row 0 has the extra bin 200.
    {
      c2* pp = sm.P + r * 16 + p;
      if (w != 0) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) pp[k2 * 320] = power2(o[k2], shfl16(o[19 - k2]));
      } else {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
          const c2 c = h ? o[19 - k2] : o[(20 - k2) % 20];
          pp[k2 * 320] = power2(o[k2], c);
        }
        if (h == 0) pp[10 * 320] = power2(o[10], o[10]);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");    // the next group's PCM (requested after barrier B)
    __syncthreads();                                   // (E)
    // ---- mel rows of this half-warp: filterbank, log10, store, running maximum.  The two half-warps of a warp walk
    // rows padded to the same number of quads (zero weights over finite padding), so the loops carry no guards.
    {
      const int fa = f0 + p;
      const bool oka = fa < n_frames, okb = fa + 16 < n_frames;
      float* dst = out + (static_cast<long long>(out_clip) * NMELS) * n_frames + fa;
      const uint32_t pbase = smem_u32(sm.P + p), wbase = smem_u32(sm.w4);
      constexpr int NS = (NMELS + 19) / 20;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int4 d = sm.slot[w][s][h];
        uint32_t qa = pbase + d.x, wq = wbase + d.y;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 1
        for (int j = 0; j < d.z; ++j) {
          float c0, c1, c2v, c3;
          float2 q0, q1, q2, q3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c0), "=f"(c1), "=f"(c2v), "=f"(c3) : "r"(wq));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(q0.x), "=f"(q0.y) : "r"(qa));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + 128];" : "=f"(q1.x), "=f"(q1.y) : "r"(qa));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + 256];" : "=f"(q2.x), "=f"(q2.y) : "r"(qa));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + 384];" : "=f"(q3.x), "=f"(q3.y) : "r"(qa));
          a0 = fmaf(c0, q0.x, a0);  b0 = fmaf(c0, q0.y, b0);
          a1 = fmaf(c1, q1.x, a1);  b1 = fmaf(c1, q1.y, b1);
          a0 = fmaf(c2v, q2.x, a0); b0 = fmaf(c2v, q2.y, b0);
          a1 = fmaf(c3, q3.x, a1);  b1 = fmaf(c3, q3.y, b1);
          qa += 512; wq += 16;
        }
        if (d.w >= 0) {
          const float va = log10_fast(fmaxf(a0 + a1, 1e-10f)), vb = log10_fast(fmaxf(b0 + b1, 1e-10f));
          float* o2 = dst + d.w * n_frames;
          if (oka) { o2[0] = va; run_max = fmaxf(run_max, va); }
          if (okb) { o2[16] = vb; run_max = fmaxf(run_max, vb); }
        }
      }
    }
  }
  flush();
  if (lane == 0 && all_max > -3.0e38f) atomicMax(max_keys + n_clips, float_key(all_max));
}

// out = (max(out, mx - 8) + 4) / 4 with mx = per-clip max (mode 1) or whole-tensor max (mode 0, the
// reference's semantics for batched input: audio.py:159).
__global__ void __launch_bounds__(256)
logmel_finish_kernel(float* __restrict__ out, long long per_clip, int n_clips, const int* __restrict__ max_keys,
                     int mode) {
  const int clip = gridDim.y - 1 - blockIdx.y;   // newest clips first: their rows may still be in L2
  const float mx = key_float(mode == 1 ? max_keys[clip] : max_keys[n_clips]);
  const float floor_v = mx - 8.0f;
  float* p = out + clip * per_clip;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((per_clip & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const long long n4 = per_clip >> 2;
    auto fix = [&](float4 v) {
      v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
      v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
      v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
      v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
      return v;
    };
    for (; i + 3 * stride < n4; i += 4 * stride) {   // four independent 16-byte loads in flight per thread
      const float4 v0 = p4[i], v1 = p4[i + stride], v2 = p4[i + 2 * stride], v3 = p4[i + 3 * stride];
      p4[i] = fix(v0);
      p4[i + stride] = fix(v1);
      p4[i + 2 * stride] = fix(v2);
      p4[i + 3 * stride] = fix(v3);
    }
    for (; i < n4; i += stride) p4[i] = fix(p4[i]);
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
  const int smem2 = static_cast<int>(sizeof(MelSmem2<NMELS>));
  const int smem3 = static_cast<int>(sizeof(MelSmem3<NMELS>));
  static PerDeviceOnce configured;  // function attributes are per device
  if (configured.first_use()) {
    WF_CHECK_CUDA(cudaFuncSetAttribute(logmel_fft_kernel<NMELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    WF_CHECK_CUDA(cudaFuncSetAttribute(logmel_fft2_kernel<NMELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    WF_CHECK_CUDA(cudaFuncSetAttribute(logmel_fft3_kernel<NMELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
  }
  WF_CHECK_CUDA(cudaMemsetAsync(keys, 0x80, (n_clips + 1) * sizeof(int), stream));
  static const bool first_layout = getenv("WF_LOGMEL_V1") != nullptr;   // A/B runs only
  static const bool second_layout = getenv("WF_LOGMEL_V2") != nullptr;
  if (!first_layout && !second_layout) {
    const int groups3 = (n_frames + GF3 - 1) / GF3;
    const long long total3 = static_cast<long long>(n_clips) * groups3;
    const int grid = static_cast<int>(total3 < 2LL * num_sms() ? total3 : 2LL * num_sms());
    logmel_fft3_kernel<NMELS><<<grid, TH3, smem3, stream>>>(pcm, clip_stride, n_samples, n_frames, n_clips, groups3,
                                                           out, keys);
  } else if (first_layout) {
    const long long max_grid = 2LL * num_sms();
    const int grid = static_cast<int>(total < max_grid ? total : max_grid);
    logmel_fft_kernel<NMELS><<<grid, THREADS, smem, stream>>>(pcm, clip_stride, n_samples, n_frames, n_clips, groups,
                                                             out, keys);
  } else {
    const int groups2 = (n_frames + GF2 - 1) / GF2;
    const long long total2 = static_cast<long long>(n_clips) * groups2;
    const int grid = static_cast<int>(total2 < num_sms() ? total2 : num_sms());
    logmel_fft2_kernel<NMELS><<<grid, TH2, smem2, stream>>>(pcm, clip_stride, n_samples, n_frames, n_clips, groups2,
                                                           out, keys);
  }
  WF_CHECK_LAUNCH();
  const long long per_clip = static_cast<long long>(NMELS) * n_frames;
  long long bx = (per_clip / 4 + 255) / 256;
  const long long want = (8LL * num_sms() + n_clips - 1) / n_clips;
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
