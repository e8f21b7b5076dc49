/* libwf - C ABI of the B200-native Whisper-Flamingo inference hot path.
 *
 * The reference (jerryyang1231/whisper-flamingo) is pure Python and owns no FFI/plugin layer
 * (SURVEY.md section 8b): its hot path sits behind the `whisper/` module API and runs as PyTorch
 * ATen library calls.  This header is therefore the *proposed* boundary: each entry point names the
 * reference code it replaces (file:line relative to the reference root).  The Python drop-in
 * package (`whisper-flamingo_b200/whisper`) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. the PyTorch allocator) unless its
 *     name ends in `_host`; the library never allocates device memory and never synchronises;
 *   - every call enqueues work on `stream` (a cudaStream_t) and is CUDA-graph capturable;
 *   - return value: 0 = ok, <0 = error (WF_ERR_*), message via wf_last_error() (thread-local);
 *   - dtype: WF_F32 (token-exact fp32 engine, CUDA cores) or WF_BF16 (tcgen05 tensor-core engine);
 *   - matrices are row-major, `ld*` are row strides in ELEMENTS.
 */
#ifndef WF_H_
#define WF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WF_OK 0
#define WF_ERR_INVALID (-1)
#define WF_ERR_CUDA (-2)
#define WF_ERR_UNSUPPORTED (-3)

#define WF_F32 0
#define WF_BF16 1

#define WF_ACT_NONE 0
#define WF_ACT_GELU 1 /* exact erf GELU (nn.GELU default, model.py:151,167,239-240) */

#define WF_LOGMEL_GLOBAL_MAX 0   /* audio.py:159 semantics: clamp with the max over the whole batch */
#define WF_LOGMEL_PER_CLIP_MAX 1 /* engine semantics: each clip clamped with its own max              */

typedef void* wf_stream_t; /* cudaStream_t */

int wf_version(void);
const char* wf_last_error(void);
/* Number of SMs of the current device (148 on B200); also forces context/attribute initialisation. */
int wf_device_sms(void);
/* Number of kernels this library has launched (or recorded into a CUDA graph under capture) so far. */
unsigned long long wf_kernel_launch_count(void);

/* ---- log-mel frontend: whisper/audio.py:111-161 (log_mel_spectrogram) ------------------------- */
/* Uploads the [n_mels, 201] fp32 mel filterbank (audio.py:92-108) from HOST memory; n_mels in {80,128}. */
int wf_logmel_set_filters(int n_mels, const float* filters_host);
long long wf_logmel_workspace_bytes(int n_clips);
/* pcm [n_clips, n_samples] fp32 (rows clip_stride apart) -> out [n_clips, n_mels, n_samples/160] fp32. */
int wf_logmel_f32(const float* pcm, int n_clips, int n_samples, long long clip_stride, int n_mels, int mode,
                  float* out, void* workspace, wf_stream_t stream);

/* ---- linear layers: whisper/model.py:35-41 (Linear), :44-50 (Conv1d via im2col), :336-338 (logits) */
typedef struct {
  void* C;              /* [M, ldc] output */
  long long ldc;
  const float* bias;    /* [N] fp32 or NULL */
  const void* residual; /* dtype of C, or NULL; row m read at (m % res_row_mod) when res_row_mod > 0 */
  long long ldr;
  int res_row_mod;
  const float* gate;    /* device scalar g or NULL: v *= tanh(g) (model.py:132,197) */
  int act;              /* WF_ACT_* */
  int out_f32;          /* WF_BF16 only: store fp32 (logits) */
  const int* c_off_ptr; /* device int p or NULL: C += p * c_off_mul elements (KV-cache append) */
  long long c_off_mul;
  /* head-major output for K/V caches when hm_heads > 0 (then N == 64 * hm_heads, ldc ignored): element (m, n)
   * goes to (((m / hm_rpb) * hm_heads + n / 64) * hm_T + m % hm_rpb) * 64 + n % 64, so that the K (or V) rows
   * of one (batch, head) are contiguous 128-byte lines. */
  int hm_heads, hm_T, hm_rpb;
  /* optional split-K workspace (WF_BF16, M <= 256): >= 4096 bytes of ZEROED arrival counters followed by
   * fp32 partial tiles; NULL disables split-K.  Must not be shared by concurrently running wf_linear calls. */
  void* ws;
  long long ws_bytes;
  /* fused LayerNorm of the A rows (model.py:30-32 in front of a Linear; WF_BF16, M <= 128 only): when ln_colsum is
   * non-NULL, A holds the RAW rows x, W holds W * diag(gamma), ln_colsum[n] = sum_k W'[n,k] (fp32) and `bias` holds
   * bias + W beta; the kernel computes the per-row mean / rstd itself: y = rstd * (x W'^T - mean * ln_colsum) + bias. */
  const float* ln_colsum;
  float ln_eps;
  /* two-destination output (fused q | k,v projection) when split_n > 0: columns [0, split_n) are stored row-major in C
   * (ldc), columns [split_n, N) in C2 with the head-major / c_off addressing above (N - split_n == 64 * hm_heads). */
  int split_n;
  void* C2;
  /* LayerNorm statistics across GEMMs (WF_BF16, M > 128; smaller problems compute them inside the consuming kernel):
   * stat_out [M, 2 * ceil(N / tile), 2] fp32 receives per-row partial (sum, sum of squares) of the stored values;
   * a later wf_linear with ln_colsum set reads them back through stat_in / stat_in_slots (= 2 * ceil(N_producer / tile),
   * tile = the producer's tile_hint). */
  float* stat_out;
  const float* stat_in;
  int stat_in_slots;
} wf_epilogue_t;
/* C = residual + tanh(gate) * act(A[M,K] . W[N,K]^T + bias).  tile_hint: 0 = auto, else N-tile 32/64/128/256. */
int wf_linear(int dtype, const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
              const wf_epilogue_t* ep, int tile_hint, wf_stream_t stream);

/* ---- LayerNorm: whisper/model.py:30-32 (fp32 statistics, eps as given) ------------------------- */
int wf_layernorm(int dtype, const void* x, long long ldx, const float* weight, const float* bias, void* y,
                 long long ldy, int rows, int d, float eps, wf_stream_t stream);

/* ---- conv stem helper: whisper/model.py:239-240 (kernel 3, padding 1, stride 1|2) as im2col ----- */
/* in(b,c,t) = in[b*in_sb + c*in_sc + t*in_st];  out [B*T_out, 3C], column = c*3 + tap. */
int wf_im2col_k3(int in_dtype, int out_dtype, const void* in, long long in_sb, long long in_sc, long long in_st,
                 int B, int C, int T_in, int stride, void* out, wf_stream_t stream);

/* ---- token + positional embedding: whisper/model.py:306-310 ------------------------------------ */
/* out[r*n_pos + j] = tok_emb[tokens[r*tok_stride + pos + j]] + pos_emb[pos + j], j < n_pos;
 * pos = *pos_ptr if non-NULL else pos_const.
 * tok_emb / pos_emb are the fp32 master tables. */
int wf_embed(int dtype, const int* tokens, long long tok_stride, const int* pos_ptr, int pos_const, int n_pos,
             const float* tok_emb, const float* pos_emb, void* out, long long ldo, int R, int d, wf_stream_t stream);
/* out[r,:] = float(in[r,:]) + table[r % mod,:]: positional add for features / stem (model.py:250, :322) */
int wf_add_rowmod(int in_dtype, int out_dtype, const void* in, long long ldi, const float* table, void* out,
                  long long ldo, long long rows, int d, int mod, wf_stream_t stream);

int wf_cast(int in_dtype, int out_dtype, const void* in, void* out, long long n, wf_stream_t stream);

/* ---- attention over full sequences: whisper/model.py:93-108 (qkv_attention) -------------------- */
/* q rows b*Tq+t, k/v rows b*Tk+t, head h at columns [64h, 64h+64); causal: key j visible to query i iff j<=i. */
int wf_attention(int dtype, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                 long long ldv, void* o, long long ldo, int B, int Tq, int Tk, int H, int causal,
                 wf_stream_t stream);

/* ---- one-token attention over cached K/V (replaces the per-step recompute of decoding.py:155-164) */
/* q,o [R, *]; rows r = kvb*G + g share cache entry kvb (beams of one audio); K/V row j of head h of entry kvb at
 * kc + kvb*kv_batch_stride + h*kv_head_stride + j*ld_kv (row-interleaved cache: head stride 64, ld = 2d;
 * head-major cache: head stride T*64, ld = 64); length = (*len_ptr + len_add) if len_ptr else len_const
 * (len_const must then hold the maximum possible length). */
long long wf_attention_decode_workspace_bytes(int R, int H);
int wf_attention_decode(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                        long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R, int G,
                        int H, const int* len_ptr, int len_add, int len_const, void* workspace,
                        long long workspace_bytes, wf_stream_t stream);
/* Same for one query per cache entry (G = 1) with a per-position indirection: key j of entry r is read from entry
 * row_table[r * table_ld + j].  Beam search re-orders hypotheses (BeamSearchDecoder + rearrange_kv_cache,
 * decoding.py:173-180, 366-368) by rewriting this table instead of moving K/V rows. */
int wf_attention_decode_paged(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                              long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R,
                              int H, const int* len_ptr, int len_add, int len_const, const int* row_table,
                              int table_ld, void* workspace, long long workspace_bytes, wf_stream_t stream);

/* ---- one-token cross-attention over the SOURCE rows (absorbed key / value projections, WF_BF16, head_dim 64,
 *      n_state a multiple of 256 up to 1280).  Same result as wf_attention_decode over K = src Wk^T, V = src Wv^T + bv
 *      (whisper/model.py:82-108 with xa, :110-134 with xt; per-step recompute of decoding.py:155-164), but the step
 *      streams src [B, T, d] once for all heads instead of a per-layer K/V cache of twice the size:
 *        wf_latent_query      qp[r, h, :]  = Wk_h^T q[r, 64h:64h+64]          (wkT = Wk^T, [d, d] row-major bf16)
 *        wf_latent_attention  ctx[b, h, :] = softmax(src_b qp[b, h, :] / 8)^T src_b     (tcgen05 + cluster kernel)
 *        wf_latent_value      o[r, 64h:64h+64] = Wv_h ctx[r, h, :] + bv_h     (wv = value.weight rows, ld = ldw)
 *      qp, ctx: [R, H, d] bf16 contiguous; one query row per source clip (R == B). */
int wf_latent_query(const void* q, long long ldq, const void* wkT, void* qp, int R, int H, wf_stream_t stream);
int wf_latent_attention(const void* qp, const void* src, void* ctx, int B, int T, int H, wf_stream_t stream);
int wf_latent_value(const void* ctx, const void* wv, long long ldw, const float* bv, void* o, long long ldo, int R,
                    int H, wf_stream_t stream);
/* Split form for full batches (more clips than SM pairs): the attention kernel is persistent, its 2-CTA clusters take
 * equal ranges of the B x ceil(T / 64) tile sequence, and a clip cut at a range border is left as TWO contexts, each
 * normalised by its own row sum: ctx [2][B, H, d] (parts part_stride elements apart), ml [2][B][32] float pairs
 * (reference maximum in log2 units, row sum; row sum 0 = part absent).  wf_latent_value_split projects both parts and
 * blends them with the softmax weights of the two segments (the projection is linear in the context).
 * wf_latent_split_supported(H) = 1 when the one-pass kernel serves n_state = 64 H (a multiple of 256, >= 512). */
int wf_latent_split_supported(int H);
int wf_latent_attention_split(const void* qp, const void* src, void* ctx, long long part_stride, float* ml, int B, int T,
                              int H, wf_stream_t stream);
int wf_latent_value_split(const void* ctx, long long part_stride, const float* ml, const void* wv, long long ldw,
                          const float* bv, void* o, long long ldo, int R, int H, wf_stream_t stream);

/* ---- sampling: whisper/decoding.py:427-442 (SuppressBlank/SuppressTokens), :276-302 (GreedyDecoder),
 *      :697-701 (no_speech_prob), loop bookkeeping of :688-718 ------------------------------------ */
/* state (device ints): [0]=t position of the token fed this step, [1]=n_init, [2]=all_done,
 *                      [3]=rows at EOT this step (scratch), [4]=sot_index. */
typedef struct {
  const float* logits; /* [R, ld] fp32, last position */
  long long ld;
  int R, V;
  const uint8_t* suppress;       /* [V] */
  const uint8_t* suppress_first; /* [V] or NULL: extra mask for the first sampled token */
  int* tokens;                   /* [R, T_cap]; [.,0..n_init) pre-filled */
  int T_cap;
  int* state;
  float* sum_logprobs;   /* [R] */
  float* no_speech_prob; /* [R] */
  int eot;
  int no_speech;       /* id or -1 */
  int timestamp_begin; /* ApplyTimestampRules (decoding.py:445-509); < 0 disables them */
  int no_timestamps;   /* id or -1 */
  int max_initial_ts;  /* max_initial_timestamp_index or -1 */
  float temperature;   /* 0 = greedy argmax; > 0 = sample softmax(logits / T) (decoding.py:286-287) */
  unsigned long long seed; /* RNG stream for temperature > 0; XORed with the 64-bit value in state[5] | state[6] << 32 */
} wf_sample_t;
int wf_sample_greedy(const wf_sample_t* args, wf_stream_t stream);
int wf_step_advance(int* state, int R, wf_stream_t stream);
/* log_softmax(filtered logits) top-k per row: whisper/decoding.py:337-347 (BeamSearchDecoder.update) */
typedef struct {
  const float* logits; /* [R, ld] */
  long long ld;
  int R, V;
  const uint8_t* suppress;
  const uint8_t* suppress_first;
  const int* tokens; /* [R, T_cap] hypotheses so far (needed by the timestamp rules) or NULL */
  int T_cap;
  int n_init;  /* sample_begin */
  int cur_len; /* tokens per row so far */
  int eot;
  int timestamp_begin, no_timestamps, max_initial_ts;
  int k;
  float* out_vals; /* [R, k] */
  int* out_idx;    /* [R, k] */
} wf_topk_t;
int wf_topk_logprobs(const wf_topk_t* args, wf_stream_t stream);
/* One beam-search step on the device: whisper/decoding.py:337-386 (BeamSearchDecoder.update) + :173-180
 * (rearrange_kv_cache, as a per-position row table) for ALL audios of the batch, graph-replayable - the position, the
 * prompt length and the SOT index are read from `state`, nothing happens while the prompt is fed (except
 * no_speech_prob at the SOT position), and state[3] counts the audios that hold max_candidates finished sequences
 * (wf_step_advance(state, n_audio) then raises state[2]).  Top-(G + 1) of log_softmax(filtered logits) per hypothesis,
 * candidate merge per audio in descending cumulative log-probability with the reference's dictionary semantics
 * (identical hypotheses - equal hyp_id - collapse), EOT continuations -> finished lists, the first G others -> the next
 * hypotheses: tokens / row_table rows permuted and extended in place, sum_logprobs updated. */
typedef struct {
  const float* logits; /* [R, ld] */
  long long ld;
  int R, V, G;         /* R = n_audio * G */
  const uint8_t* suppress;
  const uint8_t* suppress_first;
  int* tokens;     /* [R, T_cap] */
  int* tokens_tmp; /* [R, T_cap] scratch */
  int T_cap;
  int* state;
  int eot, no_speech; /* no_speech: id or -1 */
  int timestamp_begin, no_timestamps, max_initial_ts;
  int max_candidates;     /* round(beam_size * patience), 1 .. 32 */
  float* sum_logprobs;    /* [R] */
  float* sum_scratch;     /* [R] */
  float* no_speech_prob;  /* [R] */
  int* hyp_id;            /* [R]: equal ids <=> equal token sequences; start with the audio index */
  int* row_table;         /* [R, table_ld] or NULL */
  int* table_tmp;         /* [R, table_ld] scratch (NULL with row_table) */
  int table_ld;
  float* top_vals;        /* [R, G + 1] scratch */
  int* top_idx;           /* [R, G + 1] scratch */
  int* fin_tokens;        /* [n_audio, max_candidates, T_cap] */
  float* fin_score;       /* [n_audio, max_candidates] */
  int* fin_len;           /* [n_audio, max_candidates] */
  int* n_fin;             /* [n_audio], zeroed by the caller before the first step */
} wf_beam_t;
int wf_beam_step(const wf_beam_t* args, wf_stream_t stream);
/* dst row r = src row src_index[r] (first used_bytes of row_bytes): decoding.py:173-180 (rearrange_kv_cache) */
int wf_kv_gather_rows(const void* src, void* dst, const int* src_index, int R, long long row_bytes,
                      long long used_bytes, wf_stream_t stream);

/* Test hook for the temperature > 0 sampler (decoding.py:286-287 replaced by Gumbel-max over a counter RNG):
 * out_min_max[0], [1] (device floats) = smallest / largest of n uniforms drawn the way wf_sample_greedy draws them.
 * Both must lie strictly inside (0, 1), otherwise -log(-log u) is infinite and a random token wins the argmax. */
int wf_debug_uniform_range(unsigned long long seed, long long n, float* out_min_max, wf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WF_H_ */
