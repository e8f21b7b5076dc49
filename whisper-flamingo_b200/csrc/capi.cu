// extern "C" boundary of libwf (declared in include/wf.h): argument validation + dispatch only.
#include "common.cuh"
#include "kernels.h"
#include <stdarg.h>
#include <stdlib.h>

namespace wf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return WF_ERR_CUDA;
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

bool pdl_enabled(int kind) {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("WF_PDL_MASK");
    // default: GEMMs only.  Measured on B200 (large-v2 decode loop): GEMM-only PDL -5 %, attention-only -2 %,
    // every kernel +40 % (early-launched CTAs of many future kernels crowd the SMs), so the chain is kept short.
    mask = e ? atoi(e) : 0x1;
    const char* n = getenv("WF_NO_PDL");
    if (n && n[0] == '1') mask = 0;
  }
  return (mask >> kind) & 1;
}

int launch_priority(int kind) {
  static int mode = -1, least = 0, greatest = 0;
  if (mode < 0) {
    const char* e = getenv("WF_PRIO");
    mode = e ? atoi(e) : 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
  }
  if (mode == 0) return INT_MIN;
  return kind == 2 ? least : greatest;
}

int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < PerDeviceOnce::MAX_DEV ? dev : PerDeviceOnce::MAX_DEV - 1;
}

int num_sms() {
  static int sms[PerDeviceOnce::MAX_DEV] = {};
  const int dev = current_device_slot();
  if (!sms[dev]) {
    if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0)
      sms[dev] = 148;
  }
  return sms[dev];
}

static LinearEpilogue to_cpp(const wf_epilogue_t* e) {
  LinearEpilogue o;
  o.C = e->C; o.ldc = e->ldc; o.bias = e->bias; o.residual = e->residual; o.ldr = e->ldr;
  o.res_row_mod = e->res_row_mod; o.gate = e->gate; o.act = e->act; o.out_f32 = e->out_f32;
  o.c_off_ptr = e->c_off_ptr; o.c_off_mul = e->c_off_mul;
  o.hm_heads = e->hm_heads; o.hm_T = e->hm_T; o.hm_rpb = e->hm_rpb; o.ws = e->ws; o.ws_bytes = e->ws_bytes;
  o.ln_colsum = e->ln_colsum; o.ln_eps = e->ln_eps; o.split_n = e->split_n; o.C2 = e->C2;
  o.stat_out = e->stat_out; o.stat_in = e->stat_in; o.stat_in_slots = e->stat_in_slots;
  return o;
}

}  // namespace wf

using namespace wf;
#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int wf_version(void) { return 100; }
const char* wf_last_error(void) { return g_err; }
int wf_device_sms(void) { return num_sms(); }
unsigned long long wf_kernel_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int wf_logmel_set_filters(int n_mels, const float* filters_host) {
  WF_REQUIRE(filters_host != nullptr, "wf_logmel_set_filters: null filters");
  return logmel_set_filters(n_mels, filters_host);
}
long long wf_logmel_workspace_bytes(int n_clips) { return logmel_workspace_bytes(n_clips); }
int wf_logmel_f32(const float* pcm, int n_clips, int n_samples, long long clip_stride, int n_mels, int mode,
                  float* out, void* workspace, wf_stream_t stream) {
  WF_REQUIRE(pcm && out, "wf_logmel_f32: null buffer");
  return logmel_f32(pcm, n_clips, n_samples, clip_stride, n_mels, mode, out, workspace, S(stream));
}

int wf_linear(int dtype, const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
              const wf_epilogue_t* ep, int tile_hint, wf_stream_t stream) {
  WF_REQUIRE(A && W && ep && ep->C, "wf_linear: null buffer");
  WF_REQUIRE(ep->act == WF_ACT_NONE || ep->act == WF_ACT_GELU, "wf_linear: unknown activation %d", ep->act);
  WF_REQUIRE(!ep->residual || ep->ldr > 0, "wf_linear: residual without a row stride");
  const LinearEpilogue e = to_cpp(ep);
  WF_REQUIRE(dtype == WF_BF16 || (!e.ln_colsum && e.split_n == 0 && !e.stat_out),
             "wf_linear: fused LayerNorm / two-destination output exist for WF_BF16 only");
  if (dtype == WF_BF16) return linear_bf16_tc(A, lda, W, ldw, M, N, K, e, tile_hint, S(stream));
  if (dtype == WF_F32) {
    WF_REQUIRE(!ep->out_f32, "wf_linear: out_f32 is only meaningful for WF_BF16");
    return linear_f32((const float*)A, lda, (const float*)W, ldw, M, N, K, e, S(stream));
  }
  WF_REQUIRE(false, "wf_linear: bad dtype %d", dtype);
}

int wf_layernorm(int dtype, const void* x, long long ldx, const float* weight, const float* bias, void* y,
                 long long ldy, int rows, int d, float eps, wf_stream_t stream) {
  WF_REQUIRE(x && y && weight && bias, "wf_layernorm: null buffer");
  return layernorm(dtype, x, ldx, weight, bias, y, ldy, rows, d, eps, S(stream));
}

int wf_im2col_k3(int in_dtype, int out_dtype, const void* in, long long in_sb, long long in_sc, long long in_st,
                 int B, int C, int T_in, int stride, void* out, wf_stream_t stream) {
  WF_REQUIRE(in && out, "wf_im2col_k3: null buffer");
  return im2col_k3(in_dtype, out_dtype, in, in_sb, in_sc, in_st, B, C, T_in, stride, out, S(stream));
}

int wf_embed(int dtype, const int* tokens, long long tok_stride, const int* pos_ptr, int pos_const, int n_pos,
             const float* tok_emb, const float* pos_emb, void* out, long long ldo, int R, int d, wf_stream_t stream) {
  WF_REQUIRE(tokens && tok_emb && pos_emb && out, "wf_embed: null buffer");
  return embed_tokens(dtype, tokens, tok_stride, pos_ptr, pos_const, n_pos, tok_emb, pos_emb, out, ldo, R, d,
                      S(stream));
}

int wf_add_rowmod(int in_dtype, int out_dtype, const void* in, long long ldi, const float* table, void* out,
                  long long ldo, long long rows, int d, int mod, wf_stream_t stream) {
  WF_REQUIRE(in && table && out, "wf_add_rowmod: null buffer");
  return add_rowmod(in_dtype, out_dtype, in, ldi, table, out, ldo, rows, d, mod, S(stream));
}

int wf_cast(int in_dtype, int out_dtype, const void* in, void* out, long long n, wf_stream_t stream) {
  WF_REQUIRE(in && out, "wf_cast: null buffer");
  return cast_copy(in_dtype, out_dtype, in, out, n, S(stream));
}

int wf_attention(int dtype, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                 long long ldv, void* o, long long ldo, int B, int Tq, int Tk, int H, int causal,
                 wf_stream_t stream) {
  WF_REQUIRE(q && k && v && o, "wf_attention: null buffer");
  return attention_full(dtype, q, ldq, k, ldk, v, ldv, o, ldo, B, Tq, Tk, H, causal, S(stream));
}

long long wf_attention_decode_workspace_bytes(int R, int H) { return attention_decode_workspace_bytes(R, H); }
int wf_attention_decode(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                        long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R, int G,
                        int H, const int* len_ptr, int len_add, int len_const, void* workspace,
                        long long workspace_bytes, wf_stream_t stream) {
  WF_REQUIRE(q && kc && vc && o, "wf_attention_decode: null buffer");
  return attention_decode(dtype, q, ldq, kc, vc, ld_kv, kv_batch_stride, kv_head_stride, o, ldo, R, G, H, len_ptr,
                          len_add, len_const, workspace, workspace_bytes, nullptr, 0, S(stream));
}
int wf_attention_decode_paged(int dtype, const void* q, long long ldq, const void* kc, const void* vc, long long ld_kv,
                              long long kv_batch_stride, long long kv_head_stride, void* o, long long ldo, int R,
                              int H, const int* len_ptr, int len_add, int len_const, const int* row_table,
                              int table_ld, void* workspace, long long workspace_bytes, wf_stream_t stream) {
  WF_REQUIRE(q && kc && vc && o && row_table, "wf_attention_decode_paged: null buffer");
  return attention_decode(dtype, q, ldq, kc, vc, ld_kv, kv_batch_stride, kv_head_stride, o, ldo, R, 1, H, len_ptr,
                          len_add, len_const, workspace, workspace_bytes, row_table, table_ld, S(stream));
}

int wf_latent_query(const void* q, long long ldq, const void* wkT, void* qp, int R, int H, wf_stream_t stream) {
  WF_REQUIRE(q && wkT && qp && R > 0 && H > 0, "wf_latent_query: null buffer / empty problem");
  return latent_query(q, ldq, wkT, qp, R, H, S(stream));
}
int wf_latent_attention(const void* qp, const void* src, void* ctx, int B, int T, int H, wf_stream_t stream) {
  WF_REQUIRE(qp && src && ctx, "wf_latent_attention: null buffer");
  return latent_attention(qp, src, ctx, 0, nullptr, B, T, H, S(stream));
}
int wf_latent_value(const void* ctx, const void* wv, long long ldw, const float* bv, void* o, long long ldo, int R,
                    int H, wf_stream_t stream) {
  WF_REQUIRE(ctx && wv && o && R > 0 && H > 0, "wf_latent_value: null buffer / empty problem");
  return latent_value(ctx, 0, nullptr, wv, ldw, bv, o, ldo, R, H, S(stream));
}
int wf_latent_split_supported(int H) { return H > 0 && latent_pair_supported(H) ? 1 : 0; }
int wf_latent_attention_split(const void* qp, const void* src, void* ctx, long long part_stride, float* ml, int B, int T,
                              int H, wf_stream_t stream) {
  WF_REQUIRE(qp && src && ctx && ml && part_stride >= static_cast<long long>(B) * H * H * 64,
             "wf_latent_attention_split: null buffer / parts overlap");
  return latent_attention(qp, src, ctx, part_stride, ml, B, T, H, S(stream));
}
int wf_latent_value_split(const void* ctx, long long part_stride, const float* ml, const void* wv, long long ldw,
                          const float* bv, void* o, long long ldo, int R, int H, wf_stream_t stream) {
  WF_REQUIRE(ctx && ml && wv && o && R > 0 && H > 0, "wf_latent_value_split: null buffer / empty problem");
  return latent_value(ctx, part_stride, ml, wv, ldw, bv, o, ldo, R, H, S(stream));
}

int wf_sample_greedy(const wf_sample_t* a, wf_stream_t stream) {
  WF_REQUIRE(a != nullptr, "wf_sample_greedy: null args");
  SampleArgs s;
  s.logits = a->logits; s.ld = a->ld; s.R = a->R; s.V = a->V; s.suppress = a->suppress;
  s.suppress_first = a->suppress_first; s.tokens = a->tokens; s.T_cap = a->T_cap; s.state = a->state;
  s.sum_logprobs = a->sum_logprobs; s.no_speech_prob = a->no_speech_prob; s.eot = a->eot; s.no_speech = a->no_speech;
  s.timestamp_begin = a->timestamp_begin; s.no_timestamps = a->no_timestamps; s.max_initial_ts = a->max_initial_ts;
  s.temperature = a->temperature; s.seed = a->seed;
  WF_REQUIRE(s.temperature >= 0.f, "wf_sample_greedy: negative temperature");
  WF_REQUIRE(s.sum_logprobs && s.no_speech_prob, "wf_sample_greedy: null output buffer");
  return sample_greedy(s, S(stream));
}
int wf_step_advance(int* state, int R, wf_stream_t stream) {
  WF_REQUIRE(state != nullptr, "wf_step_advance: null state");
  return step_advance(state, R, S(stream));
}
int wf_topk_logprobs(const wf_topk_t* a, wf_stream_t stream) {
  WF_REQUIRE(a != nullptr, "wf_topk_logprobs: null args");
  TopkArgs t;
  t.logits = a->logits; t.ld = a->ld; t.R = a->R; t.V = a->V; t.suppress = a->suppress;
  t.suppress_first = a->suppress_first; t.tokens = a->tokens; t.T_cap = a->T_cap; t.n_init = a->n_init;
  t.cur_len = a->cur_len; t.eot = a->eot; t.timestamp_begin = a->timestamp_begin; t.no_timestamps = a->no_timestamps;
  t.max_initial_ts = a->max_initial_ts; t.k = a->k; t.out_vals = a->out_vals; t.out_idx = a->out_idx;
  return topk_logprobs(t, S(stream));
}
int wf_beam_step(const wf_beam_t* a, wf_stream_t stream) {
  WF_REQUIRE(a != nullptr, "wf_beam_step: null args");
  WF_REQUIRE(a->logits && a->suppress && a->top_vals && a->top_idx && a->state, "wf_beam_step: null buffer");
  TopkArgs t;
  t.logits = a->logits; t.ld = a->ld; t.R = a->R; t.V = a->V; t.suppress = a->suppress;
  t.suppress_first = a->suppress_first; t.tokens = a->tokens; t.T_cap = a->T_cap; t.n_init = 0; t.cur_len = 0;
  t.eot = a->eot; t.timestamp_begin = a->timestamp_begin; t.no_timestamps = a->no_timestamps;
  t.max_initial_ts = a->max_initial_ts; t.k = a->G + 1; t.out_vals = a->top_vals; t.out_idx = a->top_idx;
  t.state = a->state; t.no_speech_prob = a->no_speech_prob; t.no_speech = a->no_speech;
  BeamArgs b;
  b.R = a->R; b.G = a->G; b.max_candidates = a->max_candidates; b.eot = a->eot; b.vals = a->top_vals; b.idx = a->top_idx;
  b.tokens = a->tokens; b.tokens_tmp = a->tokens_tmp; b.T_cap = a->T_cap; b.row_table = a->row_table;
  b.table_tmp = a->table_tmp; b.table_ld = a->table_ld; b.sum_logprobs = a->sum_logprobs;
  b.sum_logprobs_out = a->sum_scratch; b.hyp_id = a->hyp_id; b.state = a->state; b.fin_tokens = a->fin_tokens;
  b.fin_score = a->fin_score; b.fin_len = a->fin_len; b.n_fin = a->n_fin;
  return beam_step(t, b, S(stream));
}
int wf_kv_gather_rows(const void* src, void* dst, const int* src_index, int R, long long row_bytes,
                      long long used_bytes, wf_stream_t stream) {
  WF_REQUIRE(src && dst && src_index, "wf_kv_gather_rows: null buffer");
  return kv_gather_rows(src, dst, src_index, R, row_bytes, used_bytes, S(stream));
}

int wf_debug_uniform_range(unsigned long long seed, long long n, float* out_min_max, wf_stream_t stream) {
  return debug_uniform_range(seed, n, out_min_max, S(stream));
}

}  // extern "C"
