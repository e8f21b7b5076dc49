// Internal C++ entry points behind the C ABI of include/wf.h (one per kernel family).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wf {

struct LinearEpilogue {
  void* C = nullptr;            // [M, ldc] output (bf16/fp32 = input dtype unless out_f32)
  long long ldc = 0;
  const float* bias = nullptr;  // [N] fp32 or null
  const void* residual = nullptr;  // same dtype as C, row r read at (r % res_row_mod if res_row_mod > 0)
  long long ldr = 0;
  int res_row_mod = 0;
  const float* gate = nullptr;  // device scalar g: value *= tanh(g) before the residual add
  int act = 0;                  // 0 none, 1 exact-erf GELU
  int out_f32 = 0;              // bf16 GEMM only: write fp32
  const int* c_off_ptr = nullptr;  // device int p: C += p * c_off_mul elements (KV-cache append position)
  long long c_off_mul = 0;
  // head-major output for K/V caches (hm_heads > 0): element (m, n) is stored at
  //   (((m / hm_rpb) * hm_heads + n / 64) * hm_T + m % hm_rpb) * 64 + n % 64      (ldc is ignored)
  int hm_heads = 0, hm_T = 0, hm_rpb = 0;
  // optional split-K workspace (bf16 engine, M <= 256): first 4096 bytes = zero-initialised arrival counters
  void* ws = nullptr;
  long long ws_bytes = 0;
  // fused LayerNorm of A (bf16 engine, M <= 128): see TcEpilogue in gemm_epilogue.cuh
  const float* ln_colsum = nullptr;
  float ln_eps = 0.f;
  // two-destination output: columns [0, split_n) -> C row-major, [split_n, N) -> C2 (head-major / offset addressing)
  int split_n = 0;
  void* C2 = nullptr;
  // LayerNorm statistics handed from the producing GEMM to the consuming one (M > 128): see TcEpilogue
  float* stat_out = nullptr;
  const float* stat_in = nullptr;
  int stat_in_slots = 0;
};

// gemm_tc.cu / gemm_f32.cu
int linear_bf16_tc(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                   const LinearEpilogue& e, int tile_hint, cudaStream_t stream);
int linear_f32(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K,
               const LinearEpilogue& e, cudaStream_t stream);

// logmel.cu
int logmel_set_filters(int n_mels, const float* dense_host /* [n_mels,201] */);
int logmel_f32(const float* pcm, int n_clips, int n_samples, long long clip_stride, int n_mels, int mode, float* out,
               void* workspace, cudaStream_t stream);
long long logmel_workspace_bytes(int n_clips);

// elementwise.cu
int layernorm(int dtype, const void* x, long long ldx, const float* w, const float* b, void* y, long long ldy,
              int rows, int d, float eps, cudaStream_t stream);
int im2col_k3(int in_dtype, int out_dtype, const void* in, long long in_sb, long long in_sc, long long in_st, int B,
              int C, int T_in, int stride, void* out, cudaStream_t stream);
int embed_tokens(int dtype, const int* tokens, long long tok_stride, const int* pos_ptr, int pos_const, int n_pos,
                 const float* tok_emb, const float* pos_emb, void* out, long long ldo, int R, int d, cudaStream_t stream);
int add_rowmod(int in_dtype, int out_dtype, const void* in, long long ldi, const float* table, void* out,
               long long ldo, long long rows, int d, int mod, cudaStream_t stream);
int cast_copy(int in_dtype, int out_dtype, const void* in, void* out, long long n, cudaStream_t stream);

// attention.cu
int attention_full(int dtype, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                   long long ldv, void* o, long long ldo, int B, int Tq, int Tk, int H, int causal,
                   cudaStream_t stream);

// attention_tc.cu (tcgen05 flash attention, bf16 non-causal)
int attention_full_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      void* o, long long ldo, int B, int Tq, int Tk, int H, cudaStream_t stream);

// latent.cu (one-token cross-attention over the source rows: absorbed K / V projections, bf16)
int latent_query(const void* q, long long ldq, const void* wkT, void* qp, int R, int H, cudaStream_t stream);
// ml == nullptr: ctx [B, H, d].  ml != nullptr (split form, needs latent_pair_supported(H)): ctx [2][B, H, d] with
// the parts `part_stride` elements apart and ml [2][B][32] pairs (reference maximum in log2 units, row sum).
int latent_attention(const void* qp, const void* src, void* ctx, long long part_stride, float* ml, int B, int T, int H,
                     cudaStream_t stream);
int latent_value(const void* ctx, long long part_stride, const float* ml, const void* wv, long long ldw, const float* bv,
                 void* o, long long ldo, int R, int H, cudaStream_t stream);
// latent_pair.cu (one-pass latent_attention on 2-CTA clusters; WF_ERR_UNSUPPORTED when the shape does not fit)
bool latent_pair_supported(int H);
int latent_attention_pair(const void* qp, const void* src, void* ctx, float* ml, long long part_stride, int B, int T,
                          int H, cudaStream_t stream);

// decode.cu
int attention_decode(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                     long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R, int G, int H,
                     const int* len_ptr, int len_add, int len_const, void* workspace, long long workspace_bytes, const int* row_table,
                     int table_ld, cudaStream_t stream);
long long attention_decode_workspace_bytes(int R, int H);

struct SampleArgs {
  const float* logits;   // [R, ld]
  long long ld;
  int R, V;
  const uint8_t* suppress;        // [V] 1 = forbidden always
  const uint8_t* suppress_first;  // [V] 1 = forbidden on the first sampled step (SuppressBlank) or null
  int* tokens;                    // [R, T_cap]
  int T_cap;
  int* state;                     // device ints: [0]=t [1]=n_init [2]=all_done [3]=n_eot_this_step [4]=sot_index
  float* sum_logprobs;            // [R]
  float* no_speech_prob;          // [R]
  int eot;
  int no_speech;                  // token id or -1
  // ApplyTimestampRules (decoding.py:445-509); timestamp_begin < 0 disables the rules
  int timestamp_begin;
  int no_timestamps;              // token id or -1
  int max_initial_ts;             // max_initial_timestamp_index or -1
  // temperature > 0: sample from softmax(logits / T) (GreedyDecoder with Categorical, decoding.py:286-287) by the
  // Gumbel-max trick with a counter-based RNG keyed on (seed, row, position, token)
  float temperature;
  unsigned long long seed;
};
int sample_greedy(const SampleArgs& a, cudaStream_t stream);
int step_advance(int* state, int R, cudaStream_t stream);
struct TopkArgs {
  const float* logits;
  long long ld;
  int R, V;
  const uint8_t* suppress;
  const uint8_t* suppress_first;
  const int* tokens;  // [R, T_cap] current hypotheses (timestamp rules) or null
  int T_cap;
  int n_init;         // sample_begin
  int cur_len;        // tokens per row so far (>= n_init)
  int eot;
  int timestamp_begin, no_timestamps, max_initial_ts;
  int k;
  float* out_vals;    // [R, k] log-probabilities
  int* out_idx;       // [R, k]
  // graph-replayable form (optional): position / n_init / sot_index from the device state, no_speech_prob recorded at
  // the SOT position, nothing else done while the prompt is being fed
  const int* state = nullptr;
  float* no_speech_prob = nullptr;
  int no_speech = -1;
};
int topk_logprobs(const TopkArgs& a, cudaStream_t stream);
struct BeamArgs {
  int R, G;                 // rows = audios * G
  int max_candidates;       // round(beam_size * patience)
  int eot;
  const float* vals;        // [R, G + 1] top log-probabilities of this step (TopkArgs.out_vals)
  const int* idx;           // [R, G + 1]
  int* tokens;              // [R, T_cap] token histories, permuted and extended in place
  int* tokens_tmp;          // [R, T_cap] scratch
  int T_cap;
  int* row_table;           // [R, table_ld] physical row of every cached self-attention position, or null
  int* table_tmp;
  int table_ld;
  float* sum_logprobs;      // [R] cumulative log-probabilities
  float* sum_logprobs_out;  // [R] scratch
  int* hyp_id;              // [R] equal ids <=> equal token sequences (all beams of an audio start equal)
  int* state;               // [0] position, [1] n_init, [3] += audios with max_candidates finished sequences
  int* fin_tokens;          // [audios, max_candidates, T_cap]
  float* fin_score;         // [audios, max_candidates]
  int* fin_len;             // [audios, max_candidates]
  int* n_fin;               // [audios]
};
int beam_step(const TopkArgs& tk, const BeamArgs& b, cudaStream_t stream);
int kv_gather_rows(const void* src, void* dst, const int* src_index, int R, long long row_bytes, long long used_bytes,
                   cudaStream_t stream);
// test hook: out_min_max[0..1] (device) = min / max of n draws of the sampling RNG's uniform
int debug_uniform_range(unsigned long long seed, long long n, float* out_min_max, cudaStream_t stream);

}  // namespace wf
