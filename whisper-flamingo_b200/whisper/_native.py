"""ctypes binding of ``libwf.so`` (C ABI in ``include/wf.h``).

There is no fallback: if the shared library is missing, or a call is made without a
CUDA device, this module raises.  Torch is used only for device memory and streams -
every function here takes torch CUDA tensors, checks them and forwards raw pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

WF_F32, WF_BF16 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
LOGMEL_GLOBAL_MAX, LOGMEL_PER_CLIP_MAX = 0, 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib", "libwf.so")
_lib = None


class WfError(RuntimeError):
    pass


class _Epilogue(C.Structure):
    _fields_ = [("C", C.c_void_p), ("ldc", C.c_longlong), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("ldr", C.c_longlong), ("res_row_mod", C.c_int), ("gate", C.c_void_p), ("act", C.c_int),
                ("out_f32", C.c_int), ("c_off_ptr", C.c_void_p), ("c_off_mul", C.c_longlong),
                ("hm_heads", C.c_int), ("hm_T", C.c_int), ("hm_rpb", C.c_int), ("ws", C.c_void_p),
                ("ws_bytes", C.c_longlong), ("ln_colsum", C.c_void_p), ("ln_eps", C.c_float), ("split_n", C.c_int),
                ("C2", C.c_void_p), ("stat_out", C.c_void_p), ("stat_in", C.c_void_p), ("stat_in_slots", C.c_int)]


class _Sample(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("ld", C.c_longlong), ("R", C.c_int), ("V", C.c_int),
                ("suppress", C.c_void_p), ("suppress_first", C.c_void_p), ("tokens", C.c_void_p),
                ("T_cap", C.c_int), ("state", C.c_void_p), ("sum_logprobs", C.c_void_p),
                ("no_speech_prob", C.c_void_p), ("eot", C.c_int), ("no_speech", C.c_int),
                ("timestamp_begin", C.c_int), ("no_timestamps", C.c_int), ("max_initial_ts", C.c_int),
                ("temperature", C.c_float), ("seed", C.c_ulonglong)]


class _Topk(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("ld", C.c_longlong), ("R", C.c_int), ("V", C.c_int),
                ("suppress", C.c_void_p), ("suppress_first", C.c_void_p), ("tokens", C.c_void_p),
                ("T_cap", C.c_int), ("n_init", C.c_int), ("cur_len", C.c_int), ("eot", C.c_int),
                ("timestamp_begin", C.c_int), ("no_timestamps", C.c_int), ("max_initial_ts", C.c_int),
                ("k", C.c_int), ("out_vals", C.c_void_p), ("out_idx", C.c_void_p)]


class _Beam(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("ld", C.c_longlong), ("R", C.c_int), ("V", C.c_int), ("G", C.c_int),
                ("suppress", C.c_void_p), ("suppress_first", C.c_void_p), ("tokens", C.c_void_p),
                ("tokens_tmp", C.c_void_p), ("T_cap", C.c_int), ("state", C.c_void_p), ("eot", C.c_int),
                ("no_speech", C.c_int), ("timestamp_begin", C.c_int), ("no_timestamps", C.c_int),
                ("max_initial_ts", C.c_int), ("max_candidates", C.c_int), ("sum_logprobs", C.c_void_p),
                ("sum_scratch", C.c_void_p), ("no_speech_prob", C.c_void_p), ("hyp_id", C.c_void_p),
                ("row_table", C.c_void_p), ("table_tmp", C.c_void_p), ("table_ld", C.c_int),
                ("top_vals", C.c_void_p), ("top_idx", C.c_void_p), ("fin_tokens", C.c_void_p),
                ("fin_score", C.c_void_p), ("fin_len", C.c_void_p), ("n_fin", C.c_void_p)]


_SIGNATURES = {
    "wf_version": (C.c_int, []),
    "wf_last_error": (C.c_char_p, []),
    "wf_device_sms": (C.c_int, []),
    "wf_kernel_launch_count": (C.c_ulonglong, []),
    "wf_logmel_set_filters": (C.c_int, [C.c_int, C.c_void_p]),
    "wf_logmel_workspace_bytes": (C.c_longlong, [C.c_int]),
    "wf_logmel_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "wf_linear": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                            C.POINTER(_Epilogue), C.c_int, C.c_void_p]),
    "wf_layernorm": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                               C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "wf_im2col_k3": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wf_embed": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "wf_add_rowmod": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "wf_cast": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "wf_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                               C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wf_attention_decode_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int]),
    "wf_attention_decode": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                      C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "wf_attention_decode_paged": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                            C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_longlong, C.c_void_p]),
    "wf_latent_query": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wf_latent_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wf_latent_value": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                  C.c_int, C.c_int, C.c_void_p]),
    "wf_latent_split_supported": (C.c_int, [C.c_int]),
    "wf_latent_attention_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p]),
    "wf_latent_value_split": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                       C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "wf_sample_greedy": (C.c_int, [C.POINTER(_Sample), C.c_void_p]),
    "wf_step_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "wf_topk_logprobs": (C.c_int, [C.POINTER(_Topk), C.c_void_p]),
    "wf_beam_step": (C.c_int, [C.POINTER(_Beam), C.c_void_p]),
    "wf_kv_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                    C.c_void_p]),
    "wf_debug_uniform_range": (C.c_int, [C.c_ulonglong, C.c_longlong, C.c_void_p, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> str:
    # WF_LIB: another build of the same library (A/B runs of kernel variants); the default is the in-tree lib/libwf.so
    return os.path.normpath(os.environ.get("WF_LIB") or _LIB_PATH)


def load() -> C.CDLL:
    """Load libwf.so (no compute).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise WfError(f"libwf.so not found at {path}: build it with `python __graft_entry__.py build` "
                          f"(there is no CPU or PyTorch fallback for the hot path)")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise WfError(f"libwf error {rc}: {load().wf_last_error().decode()}")


_replayed_launches = 0


def note_graph_replay(kernels_in_graph: int) -> None:
    """A CUDA-graph replay launches kernels that never pass through the C entry points again."""
    global _replayed_launches
    _replayed_launches += kernels_in_graph


def kernel_launch_count() -> int:
    """Kernels launched by libwf so far: direct launches + kernels inside replayed CUDA graphs."""
    return int(load().wf_kernel_launch_count()) + _replayed_launches


# Optional per-call profiler used by bench.py: when set, every wrapper below brackets its launch with two
# CUDA events on the launching stream and reports (family, work, start_event, end_event).
_profiler = None


def set_profiler(fn) -> None:
    global _profiler
    _profiler = fn


class _Prof:
    __slots__ = ("family", "work", "e0")

    def __init__(self, family: str, **work):
        self.family, self.work, self.e0 = family, work, None

    def __enter__(self):
        if _profiler is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profiler(self.family, self.work, self.e0, e1)
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    if not torch.cuda.is_available():
        raise WfError("the Whisper-Flamingo B200 engine needs a CUDA device (no CPU fallback exists)")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise WfError(f"expected a CUDA tensor, got device {t.device}")


def dtype_id(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return WF_F32
    if dt == torch.bfloat16:
        return WF_BF16
    raise WfError(f"unsupported engine dtype {dt} (float32 or bfloat16)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _row_stride(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, f"need a row-major 2-D view, got strides {t.stride()}"
    return t.stride(0)


# ----------------------------------------------------------------------------- wrappers
_filters_loaded = set()


def logmel_set_filters(n_mels: int, filters: torch.Tensor) -> None:
    f = filters.detach().to("cpu", torch.float32).contiguous()
    assert f.shape == (n_mels, 201)
    _check(load().wf_logmel_set_filters(n_mels, f.data_ptr()))
    _filters_loaded.add((torch.cuda.current_device(), n_mels))


def logmel_filters_loaded(n_mels: int) -> bool:
    return (torch.cuda.current_device(), n_mels) in _filters_loaded


def logmel(pcm: torch.Tensor, n_mels: int, mode: int) -> torch.Tensor:
    """pcm [B, N] fp32 CUDA -> [B, n_mels, N // 160] fp32."""
    require_cuda(pcm)
    assert pcm.dim() == 2 and pcm.dtype == torch.float32 and pcm.stride(1) == 1
    b, n = pcm.shape
    out = torch.empty((b, n_mels, n // 160), dtype=torch.float32, device=pcm.device)
    ws = torch.empty(int(load().wf_logmel_workspace_bytes(b)), dtype=torch.uint8, device=pcm.device)
    with _Prof("logmel", bytes=b * (n * 4 + n_mels * (n // 160) * 4)):
        _check(load().wf_logmel_f32(pcm.data_ptr(), b, n, pcm.stride(0), n_mels, mode, out.data_ptr(),
                                    ws.data_ptr(), _stream()))
    return out


def linear(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, res_row_mod: int = 0, gate: Optional[torch.Tensor] = None,
           act: int = ACT_NONE, c_off_ptr: Optional[torch.Tensor] = None, c_off_mul: int = 0,
           tile_hint: int = 0, n: Optional[int] = None, head_major: Optional[tuple] = None,
           ws: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None, ln_eps: float = 1e-5,
           out2: Optional[torch.Tensor] = None, split_n: int = 0, stat_out: Optional[torch.Tensor] = None,
           stat_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = residual + tanh(gate) * act(LN?(a)[M,K] @ w[N,K]^T + bias).  All 2-D row-major views.

    ln_colsum (bf16, M <= 128): fused LayerNorm of the rows of ``a`` - ``w`` must be W * diag(gamma), ``bias`` must be
    bias + W beta and ``ln_colsum[n] = sum_k w[n, k]``.
    out2 / split_n: columns [0, split_n) go to ``out`` (row-major), the rest to the head-major cache ``out2``.
    stat_out / stat_in: fp32 ``[M, slots, 2]`` per-row partial (sum, sum of squares) written by the GEMM that produces
    a residual stream and read by the ln_colsum GEMM that consumes it (M > 128; ``slots = 2 * ceil(N / tile)``).

    head_major=(heads, T, rows_per_batch): ``out`` is a K/V cache ``[batch, heads, T, 64]`` (any view of its
    storage) and element (m, n) lands at [m // rpb, n // 64, m % rpb (+ offset), n % 64].
    ws: optional split-K workspace (uint8, first 4096 bytes zeroed once)."""
    m, k = a.shape
    n_w, k_w = w.shape
    n = n_w if n is None else n
    assert k == k_w and a.dtype == w.dtype, (a.shape, w.shape, a.dtype, w.dtype)
    dt = dtype_id(a.dtype)
    out_f32 = int(dt == WF_BF16 and out.dtype == torch.float32)
    if not out_f32:
        assert out.dtype == a.dtype
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() >= n
    if residual is not None:
        assert residual.dtype == out.dtype
    if gate is not None:
        assert gate.dtype == torch.float32
    hm = head_major or (0, 0, 0)
    ep = _Epilogue(out.data_ptr(), 0 if (head_major and not split_n) else _row_stride(out), _ptr(bias), _ptr(residual),
                   _row_stride(residual) if residual is not None else 0, res_row_mod, _ptr(gate), act, out_f32,
                   _ptr(c_off_ptr), c_off_mul, hm[0], hm[1], hm[2], _ptr(ws),
                   0 if ws is None else ws.numel() * ws.element_size(), _ptr(ln_colsum), float(ln_eps), int(split_n),
                   _ptr(out2), _ptr(stat_out), _ptr(stat_in), 0 if stat_in is None else stat_in.shape[1])
    fam = ("gemm_tc_bf16" if m > 512 else "gemm_tc_bf16_skinny") if dt == WF_BF16 else "gemm_f32"  # <= 4 row tiles: decode GEMM
    with _Prof(fam, flops=2 * m * n * k, bytes=(m * k + n * k) * a.element_size() + m * n * out.element_size()):
        _check(load().wf_linear(dt, a.data_ptr(), _row_stride(a), w.data_ptr(), _row_stride(w), m, n, k,
                                C.byref(ep), tile_hint, _stream()))
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor,
              eps: float = 1e-5) -> torch.Tensor:
    rows, d = x.shape
    assert weight.dtype == torch.float32 and bias.dtype == torch.float32 and out.dtype == x.dtype
    with _Prof("layernorm", bytes=2 * rows * d * x.element_size()):
        _check(load().wf_layernorm(dtype_id(x.dtype), x.data_ptr(), _row_stride(x), weight.data_ptr(),
                                   bias.data_ptr(), out.data_ptr(), _row_stride(out), rows, d, eps, _stream()))
    return out


def im2col_k3(x: torch.Tensor, sb: int, sc: int, st: int, b: int, c: int, t_in: int, stride: int,
              out: torch.Tensor) -> torch.Tensor:
    with _Prof("im2col", bytes=out.numel() * out.element_size() + b * c * t_in * x.element_size()):
        _check(load().wf_im2col_k3(dtype_id(x.dtype), dtype_id(out.dtype), x.data_ptr(), sb, sc, st, b, c, t_in,
                                   stride, out.data_ptr(), _stream()))
    return out


def embed(tokens: torch.Tensor, tok_stride: int, pos_ptr: Optional[torch.Tensor], pos_const: int,
          tok_emb: torch.Tensor, pos_emb: torch.Tensor, out: torch.Tensor, n_pos: int = 1) -> torch.Tensor:
    """out [R * n_pos, d]: row r*n_pos + j = tok_emb[tokens[r, pos + j]] + pos_emb[pos + j]."""
    rows, d = out.shape
    assert rows % n_pos == 0
    assert tokens.dtype == torch.int32 and tok_emb.dtype == torch.float32 and pos_emb.dtype == torch.float32
    assert tok_emb.is_contiguous() and pos_emb.is_contiguous()
    _check(load().wf_embed(dtype_id(out.dtype), tokens.data_ptr(), tok_stride, _ptr(pos_ptr), pos_const, n_pos,
                           tok_emb.data_ptr(), pos_emb.data_ptr(), out.data_ptr(), _row_stride(out), rows // n_pos, d,
                           _stream()))
    return out


def add_rowmod(x: torch.Tensor, table: torch.Tensor, out: torch.Tensor, mod: int) -> torch.Tensor:
    """out[r] = x[r] + table[r % mod]  (x, out 2-D row-major; table fp32 [>=mod, d] contiguous)."""
    rows, d = x.shape
    assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[1] == d and table.shape[0] >= mod
    _check(load().wf_add_rowmod(dtype_id(x.dtype), dtype_id(out.dtype), x.data_ptr(), _row_stride(x),
                                table.data_ptr(), out.data_ptr(), _row_stride(out), rows, d, mod, _stream()))
    return out


def cast(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    _check(load().wf_cast(dtype_id(src.dtype), dtype_id(dst.dtype), src.data_ptr(), dst.data_ptr(), src.numel(),
                          _stream()))
    return dst


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, b: int, tq: int, tk: int,
              h: int, causal: bool) -> torch.Tensor:
    """q [B*Tq, *], k/v [B*Tk, *], out [B*Tq, *] 2-D row-major views (column h*64.. is head h)."""
    with _Prof("attention_full", flops=4 * b * h * tq * tk * 64 // (2 if causal else 1),
               bytes=(2 * b * tq + 2 * b * tk) * h * 64 * q.element_size()):
        _check(load().wf_attention(dtype_id(q.dtype), q.data_ptr(), _row_stride(q), k.data_ptr(), _row_stride(k),
                                   v.data_ptr(), _row_stride(v), out.data_ptr(), _row_stride(out), b, tq, tk, h,
                                   int(causal), _stream()))
    return out


def attention_decode_workspace_bytes(r: int, h: int) -> int:
    return int(load().wf_attention_decode_workspace_bytes(r, h))


def attention_decode(q: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor, ld_kv: int, kv_batch_stride: int,
                     kv_head_stride: int, out: torch.Tensor, g: int, h: int, len_ptr: Optional[torch.Tensor],
                     len_add: int, len_const: int, ws: Optional[torch.Tensor],
                     row_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """row_table (int32 [R, >= max length], g == 1): key j of cache entry r is read from entry row_table[r, j]."""
    r = q.shape[0]
    # algorithmic bytes: K and V rows of every (audio, head) once; dynamic lengths are reported at their bound
    dyn = len_ptr is not None  # growing self-attention cache: charged at half its bound (average over a decode)
    with _Prof("attention_decode_self" if dyn else "attention_decode",
               bytes=2 * (r // g) * h * 64 * (len_const // 2 if dyn else len_const) * q.element_size()):
        if row_table is not None:
            assert g == 1 and row_table.dtype == torch.int32 and row_table.shape[0] == r
            _check(load().wf_attention_decode_paged(
                dtype_id(q.dtype), q.data_ptr(), _row_stride(q), kc.data_ptr(), vc.data_ptr(), ld_kv, kv_batch_stride,
                kv_head_stride, out.data_ptr(), _row_stride(out), r, h, _ptr(len_ptr), len_add, len_const,
                row_table.data_ptr(), _row_stride(row_table), _ptr(ws),
                0 if ws is None else ws.numel() * ws.element_size(), _stream()))
            return out
        _check(load().wf_attention_decode(dtype_id(q.dtype), q.data_ptr(), _row_stride(q), kc.data_ptr(),
                                          vc.data_ptr(), ld_kv, kv_batch_stride, kv_head_stride, out.data_ptr(),
                                          _row_stride(out), r, g, h, _ptr(len_ptr), len_add, len_const, _ptr(ws),
                                          0 if ws is None else ws.numel() * ws.element_size(), _stream()))
    return out


def latent_query(q: torch.Tensor, wk_t: torch.Tensor, qp: torch.Tensor, h: int) -> torch.Tensor:
    """qp[r, h, :] = Wk_h^T q[r, 64h:64h+64]; wk_t = key.weight^T ([d, d] contiguous bf16), qp [R, H, d] bf16."""
    r, d = q.shape[0], h * 64
    assert q.dtype == wk_t.dtype == qp.dtype == torch.bfloat16 and wk_t.is_contiguous() and qp.is_contiguous()
    assert wk_t.shape == (d, d) and qp.numel() == r * h * d
    with _Prof("latent_query", bytes=2 * (d * d + r * d + r * h * d)):
        _check(load().wf_latent_query(q.data_ptr(), _row_stride(q), wk_t.data_ptr(), qp.data_ptr(), r, h, _stream()))
    return qp


def device_sms() -> int:
    """Streaming multiprocessors of the current CUDA device."""
    return int(load().wf_device_sms())


def latent_split_supported(h: int) -> bool:
    """True when the one-pass pair kernel (csrc/latent_pair.cu) serves n_state = 64 h: the split form below exists."""
    return bool(load().wf_latent_split_supported(int(h)))


def latent_attention(qp: torch.Tensor, src: torch.Tensor, ctx: torch.Tensor, h: int,
                     ml: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ctx[b, h, :] = softmax(src[b] qp[b, h, :] / 8)^T src[b]; src [B, T, d] contiguous bf16, qp / ctx [B, H, d].
    With ``ml`` (fp32 [2, B, 32, 2]) the split form: ctx is [2, B, H, d], two separately normalised partial contexts per
    clip (the persistent kernel cuts a clip where a cluster's tile range ends) and ml their (reference maximum in log2
    units, row sum) per head; ``latent_value(..., ml=ml)`` blends them."""
    b, t, d = src.shape
    assert src.dtype == qp.dtype == ctx.dtype == torch.bfloat16 and d == 64 * h
    assert src.is_contiguous() and qp.is_contiguous() and ctx.is_contiguous() and qp.numel() == b * d * h
    # algorithmic bytes: every source row once for all heads
    # (profiled per source length: the cross-attention over 1500 encoder rows and the x-attention over 750 feature rows
    # are different launches of the same kernel)
    with _Prof(f"latent_attention_T{t}", bytes=2 * b * t * d):
        if ml is None:
            assert ctx.numel() == b * d * h
            _check(load().wf_latent_attention(qp.data_ptr(), src.data_ptr(), ctx.data_ptr(), b, t, h, _stream()))
        else:
            assert ctx.numel() == 2 * b * d * h and ml.dtype == torch.float32 and ml.is_contiguous()
            assert ml.numel() == 2 * b * 32 * 2
            _check(load().wf_latent_attention_split(qp.data_ptr(), src.data_ptr(), ctx.data_ptr(), b * d * h,
                                                    ml.data_ptr(), b, t, h, _stream()))
    return ctx


def latent_value(ctx: torch.Tensor, wv: torch.Tensor, bv: Optional[torch.Tensor], out: torch.Tensor,
                 h: int, ml: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r, 64h:64h+64] = Wv_h ctx[r, h, :] + bv_h; wv = value.weight ([d, d] rows, bf16), bv fp32 [d] or None.
    With ``ml``: ctx [2, R, H, d] / ml [2, R, 32, 2] as left by ``latent_attention(..., ml=ml)``."""
    r, d = out.shape[0], h * 64
    assert ctx.dtype == wv.dtype == out.dtype == torch.bfloat16 and ctx.is_contiguous()
    assert bv is None or bv.dtype == torch.float32
    with _Prof("latent_value", bytes=2 * (d * d + r * d + r * h * d)):
        if ml is None:
            assert ctx.numel() == r * h * d
            _check(load().wf_latent_value(ctx.data_ptr(), wv.data_ptr(), _row_stride(wv), _ptr(bv), out.data_ptr(),
                                          _row_stride(out), r, h, _stream()))
        else:
            assert ctx.numel() == 2 * r * h * d and ml.dtype == torch.float32 and ml.numel() == 2 * r * 32 * 2
            _check(load().wf_latent_value_split(ctx.data_ptr(), r * h * d, ml.data_ptr(), wv.data_ptr(), _row_stride(wv),
                                                _ptr(bv), out.data_ptr(), _row_stride(out), r, h, _stream()))
    return out


def sample_greedy(logits: torch.Tensor, v: int, suppress: torch.Tensor, suppress_first: Optional[torch.Tensor],
                  tokens: torch.Tensor, state: torch.Tensor, sum_logprobs: torch.Tensor,
                  no_speech_prob: torch.Tensor, eot: int, no_speech: int, ts=(-1, -1, -1),
                  temperature: float = 0.0, seed: int = 0) -> None:
    """ts = (timestamp_begin | -1, no_timestamps | -1, max_initial_timestamp_index | -1)."""
    assert logits.dtype == torch.float32 and tokens.dtype == torch.int32 and state.dtype == torch.int32
    assert suppress.dtype == torch.uint8 and suppress.numel() >= v
    a = _Sample(logits.data_ptr(), _row_stride(logits), logits.shape[0], v, suppress.data_ptr(),
                _ptr(suppress_first), tokens.data_ptr(), tokens.shape[1], state.data_ptr(),
                sum_logprobs.data_ptr(), no_speech_prob.data_ptr(), eot, no_speech, ts[0], ts[1], ts[2],
                float(temperature), int(seed) & 0xFFFFFFFFFFFFFFFF)
    with _Prof("sample_greedy", bytes=2 * logits.shape[0] * v * 4):
        _check(load().wf_sample_greedy(C.byref(a), _stream()))


def step_advance(state: torch.Tensor, r: int) -> None:
    _check(load().wf_step_advance(state.data_ptr(), r, _stream()))


def topk_logprobs(logits: torch.Tensor, v: int, suppress: torch.Tensor, suppress_first: Optional[torch.Tensor],
                  tokens: Optional[torch.Tensor], n_init: int, cur_len: int, eot: int, ts, k: int,
                  out_vals: torch.Tensor, out_idx: torch.Tensor) -> None:
    a = _Topk(logits.data_ptr(), _row_stride(logits), logits.shape[0], v, suppress.data_ptr(), _ptr(suppress_first),
              _ptr(tokens), 0 if tokens is None else tokens.shape[1], n_init, cur_len, eot, ts[0], ts[1], ts[2], k,
              out_vals.data_ptr(), out_idx.data_ptr())
    _check(load().wf_topk_logprobs(C.byref(a), _stream()))


def beam_step(logits: torch.Tensor, v: int, g: int, suppress: torch.Tensor, suppress_first: Optional[torch.Tensor],
              tokens: torch.Tensor, tokens_tmp: torch.Tensor, state: torch.Tensor, eot: int, no_speech: int, ts,
              max_candidates: int, sum_logprobs: torch.Tensor, sum_scratch: torch.Tensor, no_speech_prob: torch.Tensor,
              hyp_id: torch.Tensor, row_table: Optional[torch.Tensor], table_tmp: Optional[torch.Tensor],
              top_vals: torch.Tensor, top_idx: torch.Tensor, fin_tokens: torch.Tensor, fin_score: torch.Tensor,
              fin_len: torch.Tensor, n_fin: torch.Tensor) -> None:
    """One step of BeamSearchDecoder.update + rearrange_kv_cache for every audio of the batch, on the device."""
    assert tokens.shape == tokens_tmp.shape and tokens.dtype == tokens_tmp.dtype == torch.int32
    assert fin_tokens.shape[-1] == tokens.shape[1] and hyp_id.dtype == torch.int32
    a = _Beam(logits.data_ptr(), _row_stride(logits), logits.shape[0], v, g, suppress.data_ptr(), _ptr(suppress_first),
              tokens.data_ptr(), tokens_tmp.data_ptr(), tokens.shape[1], state.data_ptr(), eot, no_speech, ts[0], ts[1],
              ts[2], max_candidates, sum_logprobs.data_ptr(), sum_scratch.data_ptr(), no_speech_prob.data_ptr(),
              hyp_id.data_ptr(), _ptr(row_table), _ptr(table_tmp), 0 if row_table is None else row_table.shape[1],
              top_vals.data_ptr(), top_idx.data_ptr(), fin_tokens.data_ptr(), fin_score.data_ptr(), fin_len.data_ptr(),
              n_fin.data_ptr())
    _check(load().wf_beam_step(C.byref(a), _stream()))


def debug_uniform_range(seed: int, n: int, device) -> tuple:
    """(min, max) of n uniforms of the temperature sampler's RNG (test hook)."""
    out = torch.empty(2, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _check(load().wf_debug_uniform_range(int(seed) & 0xFFFFFFFFFFFFFFFF, int(n), out.data_ptr(), _stream()))
    lo, hi = out.cpu().tolist()
    return lo, hi


def kv_gather_rows(src: torch.Tensor, dst: torch.Tensor, index: torch.Tensor, r: int, row_bytes: int,
                   used_bytes: int, src_off: int = 0, dst_off: int = 0) -> None:
    assert index.dtype == torch.int32
    _check(load().wf_kv_gather_rows(src.data_ptr() + src_off, dst.data_ptr() + dst_off, index.data_ptr(), r,
                                    row_bytes, used_bytes, _stream()))
