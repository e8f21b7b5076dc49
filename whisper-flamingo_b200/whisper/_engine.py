"""Host side of the CUDA engine: weight packing, encoder / decoder passes, KV-cached decode sessions.

Everything numeric is a call into ``libwf.so`` (see ``_native.py`` / ``include/wf.h``); torch only
allocates device buffers, owns the streams and captures the decode step into a CUDA graph.

Data layout in HBM
  activations        [rows, d] row-major, rows = batch * time; bf16 (tensor-core engine) or fp32
  packed weights     [N, K] row-major ("K-major" for both GEMM operands), biases / LN affine fp32
  cross / x-attn KV  decode sessions: per layer [B, 2H, T_src, 64] head-major (K heads 0..H-1, V heads H..2H-1):
                     the keys (values) of one (audio, head) are T_src contiguous 128-byte lines, written in that
                     layout directly by the K/V projection GEMM epilogue (DRAM-page friendly streaming)
  self-attn KV       per layer [R, 2H, T_cap, 64], appended in place by the K/V projection GEMM at state[0]
  (teacher-forced passes keep the simpler [B * T_src, 2d] row-interleaved buffers)
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor, nn

from . import _native as nv

_ENGINE_DTYPES = (torch.float32, torch.bfloat16)


def _engine_dtype(dt: torch.dtype) -> torch.dtype:
    if dt in _ENGINE_DTYPES:
        return dt
    if dt == torch.float16:
        return torch.bfloat16  # half precision is served by the bf16 tensor-core engine
    raise nv.WfError(f"unsupported activation dtype {dt}")


def _to_dtype(t: Tensor, dt: torch.dtype) -> Tensor:
    """Device-side cast through wf_cast (fp32 <-> bf16); other pairs only for API glue (fp16 in/out)."""
    if t.dtype == dt:
        return t
    if {t.dtype, dt} == {torch.float32, torch.bfloat16}:
        src = t.contiguous()
        return nv.cast(src, torch.empty(src.shape, dtype=dt, device=src.device))
    return t.to(dt)  # fp16 boundary conversion only


class PhaseTimer:
    """Optional phase timing (WF_TIMING=1): CUDA events between the phases of one decode() call; the last call's
    milliseconds are kept in ``PhaseTimer.last`` (read by bench.py / tools).  Disabled -> zero overhead."""
    last: Dict[str, float] = {}

    def __init__(self):
        self.on = os.environ.get("WF_TIMING", "0") == "1"
        self.marks = []

    def mark(self, name: str):
        if self.on:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def report(self):
        if self.on and len(self.marks) > 1:
            torch.cuda.synchronize()
            PhaseTimer.last = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks, self.marks[1:])}


# ============================================================================ weight packing
def _version_key(module: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def _pack_w(w: Tensor, dt: torch.dtype) -> Tensor:
    w = w.detach()
    if w.dim() > 2:
        w = w.reshape(w.shape[0], -1)
    return _to_dtype(w.contiguous().float(), dt).contiguous()


def _f32(t: Optional[Tensor], n: Optional[int] = None, device=None) -> Tensor:
    if t is None:
        return torch.zeros(n, dtype=torch.float32, device=device)
    return t.detach().float().contiguous()


@dataclass
class _MhaPack:
    qkv_w: Tensor  # [3d, d]  (rows: query | key | value)
    qkv_b: Tensor  # [3d] fp32, key part zero (the key projection has no bias)
    o_w: Tensor
    o_b: Tensor
    d: int

    @property
    def q_w(self):
        return self.qkv_w[: self.d]

    @property
    def q_b(self):
        return self.qkv_b[: self.d]

    @property
    def kv_w(self):
        return self.qkv_w[self.d:]

    @property
    def kv_b(self):
        return self.qkv_b[self.d:]


def _pack_mha(m, dt) -> _MhaPack:
    d = m.query.weight.shape[0]
    dev = m.query.weight.device
    w = torch.cat([m.query.weight.detach(), m.key.weight.detach(), m.value.weight.detach()], dim=0)
    b = torch.cat([_f32(m.query.bias), torch.zeros(d, dtype=torch.float32, device=dev), _f32(m.value.bias)])
    return _MhaPack(_pack_w(w, dt), b.contiguous(), _pack_w(m.out.weight, dt), _f32(m.out.bias), d)


@dataclass
class _MlpPack:
    w1: Tensor
    b1: Tensor
    w2: Tensor
    b2: Tensor


def _pack_mlp(seq, dt) -> _MlpPack:
    return _MlpPack(_pack_w(seq[0].weight, dt), _f32(seq[0].bias), _pack_w(seq[2].weight, dt), _f32(seq[2].bias))


@dataclass
class _LnPack:
    w: Tensor
    b: Tensor


def _pack_ln(ln) -> _LnPack:
    return _LnPack(_f32(ln.weight), _f32(ln.bias))


@dataclass
class _BlockPack:
    attn_ln: _LnPack
    attn: _MhaPack
    mlp_ln: _LnPack
    mlp: _MlpPack
    cross_ln: Optional[_LnPack] = None
    cross: Optional[_MhaPack] = None
    x_ln: List[_LnPack] = field(default_factory=list)
    x_attn: List[_MhaPack] = field(default_factory=list)
    x_gate: List[Tensor] = field(default_factory=list)
    ff_ln: Optional[_LnPack] = None
    ff: Optional[_MlpPack] = None
    ff_gate: Optional[Tensor] = None


def _pack_block(blk, dt) -> _BlockPack:
    bp = _BlockPack(_pack_ln(blk.attn_ln), _pack_mha(blk.attn, dt), _pack_ln(blk.mlp_ln), _pack_mlp(blk.mlp, dt))
    if blk.cross_attn is not None:
        bp.cross_ln, bp.cross = _pack_ln(blk.cross_attn_ln), _pack_mha(blk.cross_attn, dt)
    if blk.add_gated_x_attn != 0:
        for sub in blk.gated_x_attn_layers:
            bp.x_ln.append(_pack_ln(sub.attn_ln))
            bp.x_attn.append(_pack_mha(sub.attn, dt))
            bp.x_gate.append(_f32(sub.attn_gate))
        bp.ff_ln, bp.ff, bp.ff_gate = _pack_ln(blk.ff_ln), _pack_mlp(blk.ff, dt), _f32(blk.ff_gate)
    return bp


@dataclass
class _EncoderPack:
    conv1_w: Tensor
    conv1_b: Tensor
    conv2_w: Tensor
    conv2_b: Tensor
    pos: Tensor      # [n_ctx, d] activation dtype (fused into the conv2 epilogue)
    pos_f32: Tensor  # fp32 copy for the unfused diagnostic path
    blocks: List[_BlockPack]
    ln_post: _LnPack
    d: int
    n_head: int


@dataclass
class _DecoderPack:
    tok_emb: Tensor      # fp32 master [V, d] (embedding lookup)
    pos_emb: Tensor      # fp32 master [n_ctx, d]
    tok_emb_t: Tensor    # activation dtype copy for the logits GEMM
    pos_emb_t: Tensor    # activation dtype copy (feature positional add)
    xt_w: Optional[Tensor]
    xt_b: Optional[Tensor]
    blocks: List[_BlockPack]
    ln: _LnPack
    d: int
    n_head: int
    n_ctx: int
    n_vocab: int


def _cached_pack(module: nn.Module, dt: torch.dtype, builder):
    nv.require_cuda(next(module.parameters()))
    key = (dt, _version_key(module))
    cache = module.__dict__.setdefault("_wf_pack_cache", {})
    hit = cache.get(dt)
    if hit is not None and hit[0] == key:
        return hit[1]
    pack = builder(module, dt)
    cache[dt] = (key, pack)
    return pack


def _build_encoder_pack(enc, dt) -> _EncoderPack:
    d = enc.conv1.weight.shape[0]
    pos = enc.positional_embedding.detach().float().contiguous()
    return _EncoderPack(_pack_w(enc.conv1.weight, dt), _f32(enc.conv1.bias), _pack_w(enc.conv2.weight, dt),
                        _f32(enc.conv2.bias), _to_dtype(pos, dt), pos,
                        [_pack_block(b, dt) for b in enc.blocks], _pack_ln(enc.ln_post), d, enc.n_head)


def _build_decoder_pack(dec, dt) -> _DecoderPack:
    d = dec.token_embedding.weight.shape[1]
    tok = dec.token_embedding.weight.detach().float().contiguous()
    pos = dec.positional_embedding.detach().float().contiguous()
    xt_w = xt_b = None
    if isinstance(dec.xt_projection, nn.Linear):
        xt_w, xt_b = _pack_w(dec.xt_projection.weight, dt), _f32(dec.xt_projection.bias)
    return _DecoderPack(tok, pos, _to_dtype(tok, dt), _to_dtype(pos, dt), xt_w, xt_b,
                        [_pack_block(b, dt) for b in dec.blocks], _pack_ln(dec.ln), d, dec.n_head, pos.shape[0],
                        tok.shape[0])


@dataclass
class _LnLinear:
    """LayerNorm folded into the Linear that consumes it (decode step, bf16, <= 128 rows): the GEMM kernel reads the raw
    rows and computes mean / rstd itself; y = rstd * (x W'^T - mean * colsum) + bias'  with  W' = W diag(gamma)."""
    w: Tensor        # [N, K] bf16 = bf16(W * gamma)
    colsum: Tensor   # [N] fp32 = sum_k w[n, k] (of the ROUNDED weights the kernel multiplies)
    bias: Tensor     # [N] fp32 = b + W beta
    eps: float = 1e-5


def _fold_ln(ln: _LnPack, w_master: Tensor, b: Optional[Tensor], dt) -> _LnLinear:
    w32 = w_master.detach().float()
    wf = _to_dtype((w32 * ln.w[None, :]).contiguous(), dt).contiguous()
    bias = (w32 * ln.b[None, :]).sum(dim=1)  # W beta (elementwise: weight packing is plumbing, not a library GEMM)
    if b is not None:
        bias = bias + b
    return _LnLinear(wf, wf.float().sum(dim=1).contiguous(), bias.contiguous())


@dataclass
class _BlockFold:
    x_q: List[_LnLinear]
    ff1: Optional[_LnLinear]
    self_qkv: _LnLinear
    cross_q: Optional[_LnLinear]
    mlp1: _LnLinear
    cross_wk_t: Optional[Tensor] = None   # [d, d] = key.weight^T of the cross-attention (latent path, built on demand)
    x_wk_t: Optional[List[Tensor]] = None  # the same for every gated x-attention sub-block


def _fold_block(blk, bp: _BlockPack, dt) -> _BlockFold:
    d = bp.attn.d
    qkv_master = torch.cat([blk.attn.query.weight.detach(), blk.attn.key.weight.detach(),
                            blk.attn.value.weight.detach()], dim=0)
    x_q = [_fold_ln(bp.x_ln[i], sub.attn.query.weight, bp.x_attn[i].qkv_b[:d], dt)
           for i, sub in enumerate(blk.gated_x_attn_layers)] if blk.add_gated_x_attn != 0 else []
    ff1 = _fold_ln(bp.ff_ln, blk.ff[0].weight, bp.ff.b1, dt) if bp.ff is not None else None
    cross_q = (_fold_ln(bp.cross_ln, blk.cross_attn.query.weight, bp.cross.qkv_b[:d], dt)
               if bp.cross is not None else None)
    return _BlockFold(x_q, ff1, _fold_ln(bp.attn_ln, qkv_master, bp.attn.qkv_b, dt), cross_q,
                      _fold_ln(bp.mlp_ln, blk.mlp[0].weight, bp.mlp.b1, dt))


def latent_cross_enabled(rows: int, n_group: int, n_head: int, d: int) -> bool:
    """Cross-attention of a greedy bf16 decode step over the encoder rows themselves (absorbed key / value projections,
    csrc/latent.cu) instead of over a per-layer K/V cache.  The kernel is bound by the tensor pipe, the cached path by
    HBM: it wins once the batch fills the GPU (measured on B200, large-v2: 138 vs 164 us per layer at 128 clips,
    slower at 64).  WF_LATENT=1 / 0 forces it on / off."""
    env = os.environ.get("WF_LATENT", "")
    ok = n_group == 1 and n_head <= 24 and d == 64 * n_head and d % 128 == 0
    if env in ("0", "1"):
        return ok and env == "1"
    return ok and rows >= 112


def decoder_fold(dec, p: _DecoderPack, dt) -> List[_BlockFold]:
    """Folded LayerNorm + Linear weights of every decoder block, cached on the pack (rebuilt with it)."""
    fold = p.__dict__.get("_fold")
    if fold is None:
        fold = [_fold_block(blk, bp, dt) for blk, bp in zip(dec.blocks, p.blocks)]
        p.__dict__["_fold"] = fold
    return fold


@dataclass
class _EncBlockFold:
    qkv: _LnLinear
    mlp1: _LnLinear


def encoder_fold(enc, p: _EncoderPack, dt) -> List[_EncBlockFold]:
    """attn_ln folded into the q|k|v projection and mlp_ln into the first MLP projection of every encoder block."""
    fold = p.__dict__.get("_fold")
    if fold is None:
        fold = []
        for blk, bp in zip(enc.blocks, p.blocks):
            qkv_master = torch.cat([blk.attn.query.weight.detach(), blk.attn.key.weight.detach(),
                                    blk.attn.value.weight.detach()], dim=0)
            fold.append(_EncBlockFold(_fold_ln(bp.attn_ln, qkv_master, bp.attn.qkv_b, dt),
                                      _fold_ln(bp.mlp_ln, blk.mlp[0].weight, bp.mlp.b1, dt)))
        p.__dict__["_fold"] = fold
    return fold


STAT_TILE = 256  # N tile of the GEMMs that emit LayerNorm statistics (fixes the number of partials per row)


def encoder_pack(enc, dt) -> _EncoderPack:
    return _cached_pack(enc, dt, _build_encoder_pack)


def decoder_pack(dec, dt) -> _DecoderPack:
    return _cached_pack(dec, dt, _build_decoder_pack)


# ============================================================================ building blocks
def _empty(rows: int, cols: int, dt, dev) -> Tensor:
    return torch.empty((rows, cols), dtype=dt, device=dev)


def _mlp_inplace(x: Tensor, xn: Tensor, h: Tensor, ln: _LnPack, mlp: _MlpPack, gate: Optional[Tensor] = None,
                 ws: Optional[Tensor] = None):
    """x += tanh(gate) * W2 gelu(W1 LN(x) + b1) + b2   (reference model.py:149-152, 197, 214)."""
    nv.layernorm(x, ln.w, ln.b, xn)
    nv.linear(xn, mlp.w1, h, bias=mlp.b1, act=nv.ACT_GELU, ws=ws)
    nv.linear(h, mlp.w2, x, bias=mlp.b2, residual=x, gate=gate, ws=ws)


# ============================================================================ encoder
@torch.no_grad()
def encoder_forward(enc, mel: Tensor, track_norm: bool = False):
    """AudioEncoder.forward (reference model.py:234-258) on the CUDA engine."""
    nv.require_cuda(mel)
    assert mel.dim() == 3, "mel must be (batch, n_mels, n_frames)"
    out_dtype = mel.dtype
    dt = _engine_dtype(mel.dtype)
    with torch.cuda.device(mel.device):
        p = encoder_pack(enc, dt)
        if mel.dtype == torch.float16:
            mel = mel.to(torch.bfloat16)
        mel = mel.contiguous()
        B, C, Tm = mel.shape
        d, dev = p.d, mel.device
        # conv stem as im2col + GEMM with fused bias / exact GELU (model.py:239-240)
        a1 = _empty(B * Tm, 3 * C, dt, dev)
        nv.im2col_k3(mel, C * Tm, Tm, 1, B, C, Tm, 1, a1)
        h1 = _empty(B * Tm, d, dt, dev)
        nv.linear(a1, p.conv1_w, h1, bias=p.conv1_b, act=nv.ACT_GELU)
        del a1
        T2 = (Tm - 1) // 2 + 1
        a2 = _empty(B * T2, 3 * d, dt, dev)
        nv.im2col_k3(h1, Tm * d, 1, d, B, d, Tm, 2, a2)
        del h1
        x = _empty(B * T2, d, dt, dev)
        x_norm = None
        n_ctx = p.pos.shape[0]
        # bf16: the block LayerNorms run inside the GEMMs.  Every GEMM that writes the residual stream x also emits
        # per-row (sum, sum of squares) partials; the q|k|v and MLP-up GEMMs read the raw x with gamma / beta folded
        # into their weights and normalise in the epilogue.  x is never re-read by a LayerNorm kernel.
        fused = (dt == torch.bfloat16 and not track_norm and T2 <= n_ctx
                 and os.environ.get("WF_NO_LN_FUSION", "0") != "1" and os.environ.get("WF_NO_ENC_LN_FUSION", "0") != "1")
        fold = encoder_fold(enc, p, dt) if fused else None
        stats = (torch.empty((B * T2, 2 * ((d + STAT_TILE - 1) // STAT_TILE), 2), dtype=torch.float32, device=dev)
                 if fused else None)
        skw = dict(tile_hint=STAT_TILE, stat_out=stats) if fused and B * T2 > 128 else {}
        if track_norm or T2 > n_ctx:
            # diagnostics / over-long input: stem without the fused positional add, then crop + add
            nv.linear(a2, p.conv2_w, x, bias=p.conv2_b, act=nv.ACT_GELU)
            x3 = x.view(B, T2, d)
            if track_norm:
                x_norm = torch.linalg.norm(x3.float(), dim=-1).mean()
            if T2 > n_ctx:
                x3 = x3[:, :n_ctx]
                T2 = n_ctx
            xc = x3.reshape(B * T2, d).contiguous()
            x = nv.add_rowmod(xc, p.pos_f32, _empty(B * T2, d, dt, dev), T2)
        else:
            nv.linear(a2, p.conv2_w, x, bias=p.conv2_b, act=nv.ACT_GELU, residual=p.pos[:T2], res_row_mod=T2, **skw)
        del a2
        M = B * T2
        qkv, att = _empty(M, 3 * d, dt, dev), _empty(M, d, dt, dev)
        hbuf = _empty(M, 4 * d, dt, dev)
        if fused:
            ckw = dict(stat_in=stats) if M > 128 else {}
            for bp, bf in zip(p.blocks, fold):
                nv.linear(x, bf.qkv.w, qkv, bias=bf.qkv.bias, ln_colsum=bf.qkv.colsum, ln_eps=bf.qkv.eps, **ckw)
                nv.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, B, T2, T2, p.n_head, causal=False)
                nv.linear(att, bp.attn.o_w, x, bias=bp.attn.o_b, residual=x, **skw)
                nv.linear(x, bf.mlp1.w, hbuf, bias=bf.mlp1.bias, act=nv.ACT_GELU, ln_colsum=bf.mlp1.colsum,
                          ln_eps=bf.mlp1.eps, **ckw)
                nv.linear(hbuf, bp.mlp.w2, x, bias=bp.mlp.b2, residual=x, **skw)
        else:
            xn = _empty(M, d, dt, dev)
            for bp in p.blocks:
                nv.layernorm(x, bp.attn_ln.w, bp.attn_ln.b, xn)
                nv.linear(xn, bp.attn.qkv_w, qkv, bias=bp.attn.qkv_b)
                nv.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, B, T2, T2, p.n_head, causal=False)
                nv.linear(att, bp.attn.o_w, x, bias=bp.attn.o_b, residual=x)
                _mlp_inplace(x, xn, hbuf, bp.mlp_ln, bp.mlp)
        out = _empty(M, d, dt, dev)
        nv.layernorm(x, p.ln_post.w, p.ln_post.b, out)
        out = out.view(B, T2, d)
        if out_dtype != dt:
            out = out.to(out_dtype)
    return (out, x_norm) if track_norm else out


# ============================================================================ feature preparation (x-attn inputs)
def validate_features(feats: Optional[Sequence[Tensor]], n_batch: int, device, n_langs: int) -> List[Tensor]:
    """Shape / device checks of the gated x-attention inputs.  The reference fails with a broadcasting or device error
    inside ``q @ k`` (model.py:98-102) when a feature tensor does not match the audio batch; here the K/V projection
    writes into a pre-sized arena, so a mismatch must be refused before any kernel runs."""
    if feats is None:
        raise TypeError("object of type 'NoneType' has no len(): a gated x-attn model needs its feature input "
                        "(xt_list= / x_v=)")
    feats = list(feats)
    if len(feats) > n_langs:
        raise ValueError(f"Got {len(feats)} translations but only support up to {n_langs}")
    for i, f in enumerate(feats):
        if not torch.is_tensor(f) or f.dim() != 3:
            raise ValueError(f"feature tensor {i} must be (batch, T_x, width), got "
                             f"{tuple(f.shape) if torch.is_tensor(f) else type(f)}")
        if f.shape[0] != n_batch:
            raise ValueError(f"feature tensor {i} has batch {f.shape[0]} but the audio batch is {n_batch} "
                             f"(shape {tuple(f.shape)}); features are per clip and are not broadcast")
        if f.shape[1] < 1:
            raise ValueError(f"feature tensor {i} is empty (shape {tuple(f.shape)})")
        if not f.is_cuda or f.device != device:
            raise ValueError(f"feature tensor {i} lives on {f.device} but the audio features are on {device}")
        if not f.is_floating_point():
            raise ValueError(f"feature tensor {i} must be floating point, got {f.dtype}")
    return feats


@torch.no_grad()
def prepare_features(p: _DecoderPack, xt: Tensor, dt: torch.dtype) -> Tensor:
    """xt_projection (if widths differ) + learned positional embedding + cast (reference model.py:316-325).
    Returns [B * T_x, d] in the activation dtype."""
    B, Tx, w = xt.shape
    if Tx > p.n_ctx:
        raise RuntimeError(f"The size of tensor a ({Tx}) must match the size of tensor b ({p.n_ctx}) at "
                           f"non-singleton dimension 1: feature length exceeds n_text_ctx (reference model.py:322)")
    x2 = xt.reshape(B * Tx, w)
    if w != p.d:
        if p.xt_w is None:
            raise RuntimeError(f"feature width {w} != n_state {p.d} but the model has no xt_projection")
        a = _to_dtype(x2.float() if x2.dtype == torch.float16 else x2, dt).contiguous()
        out = _empty(B * Tx, p.d, dt, xt.device)
        return nv.linear(a, p.xt_w, out, bias=p.xt_b, residual=p.pos_emb_t[:Tx], res_row_mod=Tx)
    # width already n_state: only the positional add (fp32 add, one rounding like the reference)
    src = x2.contiguous()
    if src.dtype not in (torch.float32, dt):
        src = src.float()
    return nv.add_rowmod(src, p.pos_emb, _empty(B * Tx, w, dt, xt.device), Tx)


# ============================================================================ one block over full sequences
class _BlockScratch:
    """Activation buffers of the unfused block pass (teacher-forced decoder, stand-alone block calls)."""

    def __init__(self, M: int, d: int, dt, dev):
        self.xn, self.q, self.att = _empty(M, d, dt, dev), _empty(M, d, dt, dev), _empty(M, d, dt, dev)
        self.qkv, self.h = _empty(M, 3 * d, dt, dev), _empty(M, 4 * d, dt, dev)


def _block_forward(bp: _BlockPack, ws: _BlockScratch, x: Tensor, B: int, t: int, H: int, xa2: Optional[Tensor], Ta: int,
                   feats: Sequence[Tensor], causal: bool) -> Tensor:
    """ResidualAttentionBlock.forward (reference model.py:201-215) on rows x [B * t, d]; returns the new residual
    stream (x itself, updated in place, unless several feature tensors force a copy).
    Order: gated x-attn over every feature tensor from the SAME input (deltas summed, :184-199) + gated FF ->
    self-attention -> cross-attention (when xa2 is given) -> MLP."""
    d, dt, dev = x.shape[1], x.dtype, x.device
    xn, q, att, qkv, hbuf = ws.xn, ws.q, ws.att, ws.qkv, ws.h
    if bp.x_attn and feats is not None:
        if len(feats) > 0:
            acc = x if len(feats) <= 1 else x.clone()
            for i, f in enumerate(feats):
                Tx = f.shape[0] // B
                nv.layernorm(x, bp.x_ln[i].w, bp.x_ln[i].b, xn)
                nv.linear(xn, bp.x_attn[i].q_w, q, bias=bp.x_attn[i].q_b)
                kv_x = _empty(B * Tx, 2 * d, dt, dev)
                nv.linear(f, bp.x_attn[i].kv_w, kv_x, bias=bp.x_attn[i].kv_b)
                nv.attention(q, kv_x[:, :d], kv_x[:, d:], att, B, t, Tx, H, causal=False)
                nv.linear(att, bp.x_attn[i].o_w, acc, bias=bp.x_attn[i].o_b, residual=acc, gate=bp.x_gate[i])
            x = acc
        _mlp_inplace(x, xn, hbuf, bp.ff_ln, bp.ff, gate=bp.ff_gate)
    nv.layernorm(x, bp.attn_ln.w, bp.attn_ln.b, xn)
    nv.linear(xn, bp.attn.qkv_w, qkv, bias=bp.attn.qkv_b)
    nv.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, B, t, t, H, causal=causal)
    nv.linear(att, bp.attn.o_w, x, bias=bp.attn.o_b, residual=x)
    if bp.cross is not None and xa2 is not None:
        kv_a = _empty(B * Ta, 2 * d, dt, dev)
        nv.layernorm(x, bp.cross_ln.w, bp.cross_ln.b, xn)
        nv.linear(xn, bp.cross.q_w, q, bias=bp.cross.q_b)
        nv.linear(xa2, bp.cross.kv_w, kv_a, bias=bp.cross.kv_b)
        nv.attention(q, kv_a[:, :d], kv_a[:, d:], att, B, t, Ta, H, causal=False)
        nv.linear(att, bp.cross.o_w, x, bias=bp.cross.o_b, residual=x)
    _mlp_inplace(x, xn, hbuf, bp.mlp_ln, bp.mlp)
    return x


# ============================================================================ teacher-forced decoder pass
@torch.no_grad()
def decoder_forward(dec, tokens: Tensor, xa: Tensor, xt_list: Optional[Sequence[Tensor]]):
    """TextDecoder.forward over all positions (reference model.py:292-340), offset 0, no cache."""
    nv.require_cuda(tokens, xa)
    dt = _engine_dtype(xa.dtype)
    with torch.cuda.device(xa.device):
        p = decoder_pack(dec, dt)
        gated = len(p.blocks) > 0 and len(p.blocks[0].x_attn) > 0
        B, t = tokens.shape
        if xa.dim() != 3 or xa.shape[0] != B or xa.shape[2] != p.d:
            raise ValueError(f"xa must be ({B}, T, {p.d}) for tokens of shape {tuple(tokens.shape)}, got "
                             f"{tuple(xa.shape)}")
        if gated:
            xt_list = validate_features(xt_list, B, xa.device, len(p.blocks[0].x_attn))
        if t > p.n_ctx:
            raise RuntimeError(f"token length {t} exceeds n_text_ctx {p.n_ctx}")
        d, H, dev = p.d, p.n_head, xa.device
        xa2 = _to_dtype(xa, dt).contiguous().view(-1, d)
        Ta = xa.shape[1]
        feats = [prepare_features(p, xt, dt) for xt in xt_list] if (gated and xt_list is not None) else []
        tok32 = tokens.to(torch.int32).contiguous()
        M = B * t
        x = _empty(M, d, dt, dev)
        nv.embed(tok32, t, None, 0, p.tok_emb, p.pos_emb, x, n_pos=t)
        ws = _BlockScratch(M, d, dt, dev)
        for bp in p.blocks:
            x = _block_forward(bp, ws, x, B, t, H, xa2, Ta, feats, causal=True)
        xn = ws.xn
        nv.layernorm(x, p.ln.w, p.ln.b, xn)
        logits = torch.empty((M, p.n_vocab), dtype=torch.float32, device=dev)
        nv.linear(xn, p.tok_emb_t, logits)
    return logits.view(B, t, p.n_vocab)


# ============================================================================ KV-cached decode session
class DecodeSession:
    """One batch of R = B * G decoder rows stepping one token at a time over cached K/V.

    Cross-attention and gated-x-attention K/V are projected once per clip here (the reference redoes
    that at every step, decoding.py:155-164); self-attention K/V are appended by the projection GEMM.
    The whole step is captured once into a CUDA graph and replayed; the step position lives in device
    memory (``state[0]``), so replay needs no host-side parameter updates.
    """

    def __init__(self, dec, xa: Tensor, feats: Optional[Sequence[Tensor]], n_group: int, t_cap: int,
                 use_graph: bool = True):
        nv.require_cuda(xa)
        self.dt = dt = _engine_dtype(xa.dtype)
        self.dev = dev = xa.device
        self.p = p = decoder_pack(dec, dt)
        self.G = G = n_group
        self.B = B = xa.shape[0]
        self.R = R = B * G
        self.T_cap = t_cap
        d, H = p.d, p.n_head
        self.gated = len(p.blocks) > 0 and len(p.blocks[0].x_attn) > 0
        # bf16 and at most four 128-row tiles: LayerNorms run inside the GEMMs and q | k,v is a single projection
        self.fold = (decoder_fold(dec, p, dt) if (dt == torch.bfloat16 and R <= 512 and
                                                  os.environ.get("WF_NO_LN_FUSION", "0") != "1") else None)
        feats = self._check_feats(feats)
        self.Ta = Ta = xa.shape[1]
        self.Tx = [f.shape[1] for f in feats]
        L = len(p.blocks)
        # ---- per-clip K/V caches (head-major) and step buffers; contents are (re)filled by load()
        self.latent = self.fold is not None and latent_cross_enabled(R, G, H, d)
        # the gated x-attention takes the same route (its source rows = the prepared features, shared by all layers):
        # no x-attn K/V arena and no per-layer K/V projection pass at load().  WF_LATENT_X=0 keeps the cached kernel.
        self.latent_x = self.latent and self.gated and os.environ.get("WF_LATENT_X", "1") != "0"
        if self.latent:
            # no cross-attention K/V arena at all: every layer attends over the same bf16 encoder rows
            self.cross_kv = []
            self.xa_src = torch.empty((B, Ta, d), dtype=dt, device=dev)
            self.qp = torch.empty((R, H, d), dtype=dt, device=dev)
            # full batches run the persistent one-pass kernel, which leaves a clip cut at a cluster border as two
            # partial contexts + their softmax statistics (csrc/latent_pair.cu); latent_value blends them
            self.split = nv.latent_split_supported(H) and B > nv.device_sms() // 2
            self.ctx = torch.zeros((2 if self.split else 1, R, H, d), dtype=dt, device=dev)
            self.ctx_ml = torch.zeros((2, R, 32, 2), dtype=torch.float32, device=dev) if self.split else None
            for bp, bf in zip(p.blocks, self.fold):
                if bf.cross_wk_t is None:
                    bf.cross_wk_t = bp.cross.qkv_w[d:2 * d].t().contiguous()
                if self.latent_x and bf.x_wk_t is None:
                    bf.x_wk_t = [xp.qkv_w[d:2 * d].t().contiguous() for xp in bp.x_attn]
        else:
            self.cross_kv = [torch.empty((B, 2 * H, Ta, 64), dtype=dt, device=dev) for _ in range(L)]
        if self.latent_x:
            self.x_kv = [[] for _ in range(L)]
            self.x_src = [torch.empty((B, tx, d), dtype=dt, device=dev) for tx in self.Tx]
        else:
            self.x_kv = [[torch.empty((B, 2 * H, tx, 64), dtype=dt, device=dev) for tx in self.Tx] for _ in range(L)]
        self.self_kv = [torch.zeros((R, 2 * H, t_cap, 64), dtype=dt, device=dev) for _ in range(L)]
        self.gemm_ws = torch.zeros(4096 + 8 * 1024 * 1024, dtype=torch.uint8, device=dev)  # split-K counters + partials
        self.x = _empty(R, d, dt, dev)
        self.xn, self.q, self.att = _empty(R, d, dt, dev), _empty(R, d, dt, dev), _empty(R, d, dt, dev)
        self.acc = _empty(R, d, dt, dev) if len(feats) > 1 else None
        self.h = _empty(R, 4 * d, dt, dev)
        self.v_pad = (p.n_vocab + 7) // 8 * 8
        self.logits = torch.empty((R, self.v_pad), dtype=torch.float32, device=dev)
        ws_bytes = nv.attention_decode_workspace_bytes(R, H)
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)
        self.tokens = torch.zeros((R, t_cap + 1), dtype=torch.int32, device=dev)
        self.sum_logprobs = torch.zeros(R, dtype=torch.float32, device=dev)
        self.no_speech_prob = torch.full((R,), float("nan"), dtype=torch.float32, device=dev)
        self.suppress = torch.zeros(p.n_vocab, dtype=torch.uint8, device=dev)
        self.suppress_first = torch.zeros(p.n_vocab, dtype=torch.uint8, device=dev)
        # beam search: per-position indirection of the self-attention cache (key j of row r lives in physical row
        # row_table[r, j]); re-ordering hypotheses rewrites this table instead of moving K/V rows
        self.row_table = (torch.arange(R, dtype=torch.int32, device=dev)[:, None].repeat(1, t_cap).contiguous()
                          if G > 1 else None)
        self.use_graph = use_graph and os.environ.get("WF_NO_GRAPH", "0") != "1"
        self._fwd_graph: Optional[torch.cuda.CUDAGraph] = None
        self._fwd_graph_kernels = 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_kernels = 0
        self._graph_key = None
        self._sampler = None
        self.load(xa, feats)

    def _check_feats(self, feats) -> List[Tensor]:
        if not self.gated:
            return []
        return validate_features(feats, self.B, self.dev, len(self.p.blocks[0].x_attn))

    def shape_key(self):
        return (self.dt, self.dev, self.B, self.G, self.T_cap, self.Ta, tuple(self.Tx), id(self.p))

    def load(self, xa: Tensor, feats: Sequence[Tensor]):
        """Per-clip precompute into the session's caches: cross-attention K/V of the encoder output and
        x-attention K/V of the (projected, position-embedded) features, for every layer."""
        p, dt, B, H = self.p, self.dt, self.B, self.p.n_head
        xa2 = _to_dtype(xa, dt).contiguous().view(-1, p.d)
        fprep = [prepare_features(p, f, dt) for f in feats]
        if self.latent:
            self.xa_src.view(-1, p.d).copy_(xa2)
        if self.latent_x:
            for i, f in enumerate(fprep):
                self.x_src[i].view(-1, p.d).copy_(f)
        if self.latent and self.latent_x:
            return
        for l, bp in enumerate(p.blocks):
            if not self.latent:
                nv.linear(xa2, bp.cross.kv_w, self.cross_kv[l].view(-1, 64), bias=bp.cross.kv_b,
                          head_major=(2 * H, self.Ta, self.Ta))
            if not self.latent_x:
                for i, f in enumerate(fprep):
                    nv.linear(f, bp.x_attn[i].kv_w, self.x_kv[l][i].view(-1, 64), bias=bp.x_attn[i].kv_b,
                              head_major=(2 * H, self.Tx[i], self.Tx[i]))

    # -- one decoder pass for the token at position state[0]; logits of that position land in self.logits
    def _forward_token(self):
        if self.fold is not None:
            return self._forward_token_fused()
        p, d, H, G = self.p, self.p.d, self.p.n_head, self.G
        st = self.state
        gws = self.gemm_ws if self.dt == torch.bfloat16 else None
        nv.embed(self.tokens, self.tokens.shape[1], st, 0, p.tok_emb, p.pos_emb, self.x)
        x, xn, q, att, h = self.x, self.xn, self.q, self.att, self.h
        for l, bp in enumerate(p.blocks):
            if self.gated:
                multi = len(self.x_kv[l]) > 1
                acc = x
                if multi:
                    self.acc.copy_(x)
                    acc = self.acc
                for i, kvx in enumerate(self.x_kv[l]):
                    Tx = self.Tx[i]
                    nv.layernorm(x, bp.x_ln[i].w, bp.x_ln[i].b, xn)
                    nv.linear(xn, bp.x_attn[i].q_w, q, bias=bp.x_attn[i].q_b, ws=gws)
                    nv.attention_decode(q, kvx, kvx[:, H:], 64, 2 * H * Tx * 64, Tx * 64, att, G, H, None, 0, Tx,
                                        self.ws)
                    nv.linear(att, bp.x_attn[i].o_w, acc, bias=bp.x_attn[i].o_b, residual=acc, gate=bp.x_gate[i],
                              ws=gws)
                if multi:
                    x.copy_(acc)
                _mlp_inplace(x, xn, h, bp.ff_ln, bp.ff, gate=bp.ff_gate, ws=gws)
            # causal self-attention over the cache; this token's K/V are appended by the GEMM epilogue
            Tc = self.T_cap
            nv.layernorm(x, bp.attn_ln.w, bp.attn_ln.b, xn)
            nv.linear(xn, bp.attn.q_w, q, bias=bp.attn.q_b, ws=gws)
            kv = self.self_kv[l]  # [R, 2H, T_cap, 64]; row r, position state[0]: offset state[0] * 64 inside each head
            nv.linear(xn, bp.attn.kv_w, kv.view(-1, 64), bias=bp.attn.kv_b, head_major=(2 * H, Tc, 1),
                      c_off_ptr=st, c_off_mul=64, ws=gws)
            nv.attention_decode(q, kv, kv[:, H:], 64, 2 * H * Tc * 64, Tc * 64, att, 1, H, st, 1, Tc, self.ws,
                                row_table=self.row_table)
            nv.linear(att, bp.attn.o_w, x, bias=bp.attn.o_b, residual=x, ws=gws)
            # cross-attention to the encoder output (K/V cached per audio, shared by its G beams)
            nv.layernorm(x, bp.cross_ln.w, bp.cross_ln.b, xn)
            nv.linear(xn, bp.cross.q_w, q, bias=bp.cross.q_b, ws=gws)
            ckv = self.cross_kv[l]
            nv.attention_decode(q, ckv, ckv[:, H:], 64, 2 * H * self.Ta * 64, self.Ta * 64, att, G, H, None, 0,
                                self.Ta, self.ws)
            nv.linear(att, bp.cross.o_w, x, bias=bp.cross.o_b, residual=x, ws=gws)
            _mlp_inplace(x, xn, h, bp.mlp_ln, bp.mlp, ws=gws)
        nv.layernorm(x, p.ln.w, p.ln.b, xn)
        nv.linear(xn, p.tok_emb_t, self.logits, n=p.n_vocab, ws=gws)

    def _forward_token_fused(self):
        """Same pass with every block LayerNorm folded into the GEMM behind it and q | k,v as one projection:
        11 launches per block instead of 19 (reference model.py:171-215 per block)."""
        p, H, G = self.p, self.p.n_head, self.G
        st = self.state
        nv.embed(self.tokens, self.tokens.shape[1], st, 0, p.tok_emb, p.pos_emb, self.x)
        x, q, att, h = self.x, self.q, self.att, self.h

        def lnlin(f: _LnLinear, out, **kw):
            nv.linear(x, f.w, out, bias=f.bias, ln_colsum=f.colsum, ln_eps=f.eps, **kw)

        d = p.d

        def latent_attend(src, wk_t, mp: _MhaPack):
            # q' = Wk_h^T q_h -> softmax(src q' / 8)^T src -> Wv_h c_h + bv_h: the step streams src once for all heads
            nv.latent_query(q, wk_t, self.qp, H)
            if self.split:
                nv.latent_attention(self.qp, src, self.ctx, H, ml=self.ctx_ml)
                nv.latent_value(self.ctx, mp.qkv_w[2 * d:], mp.qkv_b[2 * d:], att, H, ml=self.ctx_ml)
            else:
                nv.latent_attention(self.qp, src, self.ctx[0], H)
                nv.latent_value(self.ctx[0], mp.qkv_w[2 * d:], mp.qkv_b[2 * d:], att, H)

        for l, (bp, bf) in enumerate(zip(p.blocks, self.fold)):
            if self.gated:
                multi = len(self.Tx) > 1
                acc = x
                if multi:
                    self.acc.copy_(x)
                    acc = self.acc
                for i, Tx in enumerate(self.Tx):
                    lnlin(bf.x_q[i], q)
                    if self.latent_x:
                        latent_attend(self.x_src[i], bf.x_wk_t[i], bp.x_attn[i])
                    else:
                        kvx = self.x_kv[l][i]
                        nv.attention_decode(q, kvx, kvx[:, H:], 64, 2 * H * Tx * 64, Tx * 64, att, G, H, None, 0, Tx,
                                            self.ws)
                    nv.linear(att, bp.x_attn[i].o_w, acc, bias=bp.x_attn[i].o_b, residual=acc, gate=bp.x_gate[i])
                if multi:
                    x.copy_(acc)
                lnlin(bf.ff1, h, act=nv.ACT_GELU)
                nv.linear(h, bp.ff.w2, x, bias=bp.ff.b2, residual=x, gate=bp.ff_gate)
            Tc = self.T_cap
            kv = self.self_kv[l]
            lnlin(bf.self_qkv, q, out2=kv.view(-1, 64), split_n=p.d, head_major=(2 * H, Tc, 1), c_off_ptr=st,
                  c_off_mul=64)
            nv.attention_decode(q, kv, kv[:, H:], 64, 2 * H * Tc * 64, Tc * 64, att, 1, H, st, 1, Tc, self.ws,
                                row_table=self.row_table)
            nv.linear(att, bp.attn.o_w, x, bias=bp.attn.o_b, residual=x)
            lnlin(bf.cross_q, q)
            if self.latent:
                latent_attend(self.xa_src, bf.cross_wk_t, bp.cross)
            else:
                ckv = self.cross_kv[l]
                nv.attention_decode(q, ckv, ckv[:, H:], 64, 2 * H * self.Ta * 64, self.Ta * 64, att, G, H, None, 0,
                                    self.Ta, self.ws)
            nv.linear(att, bp.cross.o_w, x, bias=bp.cross.o_b, residual=x)
            lnlin(bf.mlp1, h, act=nv.ACT_GELU)
            nv.linear(h, bp.mlp.w2, x, bias=bp.mlp.b2, residual=x)
        nv.layernorm(x, p.ln.w, p.ln.b, self.xn)
        nv.linear(self.xn, p.tok_emb_t, self.logits, n=p.n_vocab, ws=self.gemm_ws)

    # -- greedy ---------------------------------------------------------------------------------------
    def configure_greedy(self, initial_tokens: Sequence[int], sot_index: int, suppress: Tensor,
                         suppress_first: Optional[Tensor], eot: int, no_speech: int, ts: Sequence[int],
                         temperature: float = 0.0, seed: int = 0):
        n_init = len(initial_tokens)
        init = torch.tensor(list(initial_tokens), dtype=torch.int32, device=self.dev)
        self.tokens.zero_()
        self.tokens[:, :n_init] = init
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        lo, hi = seed & 0xFFFFFFFF, seed >> 32
        lo, hi = (lo - (1 << 32) if lo >= (1 << 31) else lo), (hi - (1 << 32) if hi >= (1 << 31) else hi)
        self.state.copy_(torch.tensor([0, n_init, 0, 0, sot_index, lo, hi, 0], dtype=torch.int32))
        self.sum_logprobs.zero_()
        self.no_speech_prob.fill_(float("nan"))
        self.reset_row_table()
        # the masks live in session-owned buffers so that a captured graph stays valid across decode() calls
        self.suppress.copy_(suppress)
        if suppress_first is not None:
            self.suppress_first.copy_(suppress_first)
        self._sampler = (self.suppress, self.suppress_first if suppress_first is not None else None, eot, no_speech,
                         tuple(ts), float(temperature), 0)
        # the per-call seed lives in device memory (state[5:7]) so that a new seed does not re-capture the graph
        key = (suppress_first is not None, eot, no_speech, tuple(ts), float(temperature))
        if key != self._graph_key:  # scalars baked into the captured launch changed: re-capture lazily
            self._graph, self._graph_key = None, key
        self.n_init = n_init

    def _greedy_step(self):
        self._forward_token()
        suppress, suppress_first, eot, no_speech, ts, temperature, seed = self._sampler
        nv.sample_greedy(self.logits, self.p.n_vocab, suppress, suppress_first, self.tokens, self.state,
                         self.sum_logprobs, self.no_speech_prob, eot, no_speech, ts, temperature, seed)
        nv.step_advance(self.state, self.R)

    def set_initial_rows(self, rows: Tensor):
        """Per-clip prompts (detected language tokens differ per clip): rows int32 [B, n_init]."""
        self.tokens[:, : rows.shape[1]] = rows.repeat_interleave(self.G, dim=0)

    def results(self, n_tokens: int):
        """(tokens [R, n_tokens] list, sum_logprobs [R] list, no_speech_prob [R] list) on the host."""
        return (self.tokens[:, :n_tokens].cpu().tolist(), self.sum_logprobs.cpu().tolist(),
                self.no_speech_prob.cpu().tolist())

    def _ensure_graph(self):
        if not self.use_graph or self._graph is not None:
            return
        # warm-up outside capture (lazy module init, cudaFuncSetAttribute), then restore the state
        snap = (self.state.clone(), self.tokens.clone(), self.sum_logprobs.clone(), self.no_speech_prob.clone())
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self._greedy_step()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        for dst, src in zip((self.state, self.tokens, self.sum_logprobs, self.no_speech_prob), snap):
            dst.copy_(src)
        g = torch.cuda.CUDAGraph()
        before = nv.kernel_launch_count()
        with torch.cuda.graph(g):
            self._greedy_step()
        # capture records but does not execute: state is still at t=0
        self._graph_kernels = nv.kernel_launch_count() - before
        self._graph = g

    def step(self):
        """One token for every row: graph replay (or the eager kernel sequence) on the CURRENT stream."""
        if self._graph is not None:
            self._graph.replay()
            nv.note_graph_replay(self._graph_kernels)
        else:
            self._greedy_step()

    def all_done(self) -> bool:
        return int(self.state[2].item()) != 0  # every row has emitted EOT (decoding.py:296, 713)

    def run_greedy(self, n_sample: int, check_every: int = 8) -> int:
        """Feeds the initial tokens and samples up to ``n_sample`` new ones; returns #graph launches."""
        total = self.n_init - 1 + n_sample
        assert self.n_init + n_sample <= self.tokens.shape[1]
        launches = 0
        self._ensure_graph()
        for i in range(total):
            self.step()
            launches += 1
            sampled = i - (self.n_init - 1) + 1
            if sampled > 0 and sampled % check_every == 0 and i + 1 < total:
                if self.all_done():
                    break
        return launches

    # -- beam search on the device ------------------------------------------------------------------------
    def configure_beam(self, max_candidates: int):
        """Buffers of the device-side beam search (after configure_greedy, which sets prompt / masks / state)."""
        R, B, dev = self.R, self.B, self.dev
        ld = self.tokens.shape[1]
        if getattr(self, "_beam_cap", None) != (max_candidates, ld):
            self._beam_cap = (max_candidates, ld)
            self.tokens_tmp = torch.zeros_like(self.tokens)
            self.table_tmp = torch.zeros_like(self.row_table) if self.row_table is not None else None
            self.sum_tmp = torch.zeros(R, dtype=torch.float32, device=dev)
            self.hyp_id = torch.zeros(R, dtype=torch.int32, device=dev)
            self.top_vals = torch.empty((R, self.G + 1), dtype=torch.float32, device=dev)
            self.top_idx = torch.empty((R, self.G + 1), dtype=torch.int32, device=dev)
            self.fin_tokens = torch.zeros((B, max_candidates, ld), dtype=torch.int32, device=dev)
            self.fin_score = torch.zeros((B, max_candidates), dtype=torch.float32, device=dev)
            self.fin_len = torch.zeros((B, max_candidates), dtype=torch.int32, device=dev)
            self.n_fin = torch.zeros(B, dtype=torch.int32, device=dev)
            self._beam_graph = None
        # all hypotheses of an audio start as the same sequence: equal ids
        self.hyp_id.copy_(torch.arange(R, dtype=torch.int32, device=dev) // self.G)
        self.n_fin.zero_()
        key = (self._sampler[1] is not None,) + tuple(self._sampler[2:5]) + (max_candidates,)
        if key != getattr(self, "_beam_key", None):
            self._beam_key, self._beam_graph = key, None
        self.max_candidates = max_candidates

    def _beam_step(self):
        """Decoder pass + top-(G + 1) + candidate merge + history / cache-table permutation + position advance."""
        self._forward_token()
        suppress, suppress_first, eot, no_speech, ts = self._sampler[:5]
        nv.beam_step(self.logits, self.p.n_vocab, self.G, suppress, suppress_first, self.tokens, self.tokens_tmp,
                     self.state, eot, no_speech, ts, self.max_candidates, self.sum_logprobs, self.sum_tmp,
                     self.no_speech_prob, self.hyp_id, self.row_table, self.table_tmp, self.top_vals, self.top_idx,
                     self.fin_tokens, self.fin_score, self.fin_len, self.n_fin)
        nv.step_advance(self.state, self.B)

    def run_beam(self, n_sample: int, n_ctx: int, check_every: int = 8) -> int:
        """Feeds the prompt and runs up to n_sample beam-search steps (one replay of the captured step each); the host
        only polls the completion flag (every audio has its finished candidates) every few steps."""
        total = self.n_init - 1 + n_sample
        assert self.n_init + n_sample <= self.tokens.shape[1]
        if self.use_graph and self._beam_graph is None:
            snap = [t.clone() for t in (self.state, self.tokens, self.sum_logprobs, self.no_speech_prob, self.hyp_id,
                                        self.n_fin)] + ([self.row_table.clone()] if self.row_table is not None else [])
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._beam_step()        # warm-up outside capture
            torch.cuda.current_stream(self.dev).wait_stream(side)
            for dst, src in zip((self.state, self.tokens, self.sum_logprobs, self.no_speech_prob, self.hyp_id, self.n_fin)
                                + ((self.row_table,) if self.row_table is not None else ()), snap):
                dst.copy_(src)
            g = torch.cuda.CUDAGraph()
            before = nv.kernel_launch_count()
            with torch.cuda.graph(g):
                self._beam_step()
            self._beam_graph_kernels = nv.kernel_launch_count() - before
            self._beam_graph = g
        steps = 0
        for i in range(total):
            if self._beam_graph is not None:
                self._beam_graph.replay()
                nv.note_graph_replay(self._beam_graph_kernels)
            else:
                self._beam_step()
            steps += 1
            sampled = i - (self.n_init - 1) + 1
            if sampled > 0 and self.n_init + sampled > n_ctx:
                break                      # the reference stops once a hypothesis is longer than the text context
            if sampled > 0 and sampled % check_every == 0 and i + 1 < total and self.all_done():
                break
        return steps

    # -- beam search support, host-driven (kept for WF_BEAM_HOST=1 A/B runs; see decoding.py) --------------
    def forward_at(self, pos: int):
        """Single-position pass used by the beam-search driver: sets state[0] = pos, then replays the captured
        forward graph (or launches the kernels eagerly when graphs are off)."""
        self.state[0] = pos
        if not self.use_graph:
            self._forward_token()
            return
        if self._fwd_graph is None:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._forward_token()  # warm-up outside capture; it recomputes what the replay computes again
            torch.cuda.current_stream(self.dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            before = nv.kernel_launch_count()
            with torch.cuda.graph(g):
                self._forward_token()
            self._fwd_graph_kernels = nv.kernel_launch_count() - before
            self._fwd_graph = g
        self._fwd_graph.replay()
        nv.note_graph_replay(self._fwd_graph_kernels)

    def reorder_self_kv(self, src_index: Tensor, used_positions: int):
        """Hypothesis r continues hypothesis src_index[r] (reference rearrange_kv_cache, decoding.py:173-180): its
        first `used_positions` cache positions are read where the parent read them; later positions are its own."""
        tbl = self.row_table
        new = tbl.index_select(0, src_index.long())
        if used_positions < tbl.shape[1]:
            new[:, used_positions:] = torch.arange(self.R, dtype=torch.int32, device=self.dev)[:, None]
        tbl.copy_(new)

    def reset_row_table(self):
        if self.row_table is not None:
            self.row_table.copy_(torch.arange(self.R, dtype=torch.int32, device=self.dev)[:, None]
                                 .expand(-1, self.row_table.shape[1]))


class SplitSession:
    """Greedy decode of one batch as `n_split` independent sub-batches on `n_split` CUDA streams.

    A decode step is a serial chain of ~600 kernels: the K/V-streaming attention kernels are HBM-bound, the M<=128
    GEMMs and LayerNorms between them are latency-bound (a few microseconds each, the SMs mostly idle).  Clips are
    independent, so two (or more) sub-batches stepping concurrently let one sub-batch's latency-bound chain run under
    the other's K/V streaming.  Each sub-batch is a full DecodeSession (own step buffers, K/V arena slice and CUDA
    graph); the packed weights are shared.  Beam search stays on a single session (host-driven reordering).
    """

    def __init__(self, dec, xa: Tensor, feats: Optional[Sequence[Tensor]], n_group: int, t_cap: int, n_split: int):
        B = xa.shape[0]
        n_split = max(1, min(n_split, B))
        bounds = [(B * i // n_split, B * (i + 1) // n_split) for i in range(n_split)]
        self.bounds = bounds
        self.B, self.G, self.R, self.T_cap, self.dev = B, n_group, B * n_group, t_cap, xa.device
        self.subs = [DecodeSession(dec, xa[lo:hi], None if feats is None else [f[lo:hi] for f in feats], n_group, t_cap)
                     for lo, hi in bounds]
        self.streams = [torch.cuda.Stream(device=xa.device) for _ in bounds]
        first = self.subs[0]
        self.gated, self.Tx, self.dt, self.p, self.Ta = first.gated, first.Tx, first.dt, first.p, first.Ta
        self.n_split = n_split

    def _check_feats(self, feats):
        if not self.gated:
            return []
        return validate_features(feats, self.B, self.dev, len(self.p.blocks[0].x_attn))

    def shape_key(self):
        return (self.dt, self.dev, self.B, self.G, self.T_cap, self.Ta, tuple(self.Tx), id(self.p), self.n_split)

    def load(self, xa: Tensor, feats: Sequence[Tensor]):
        for sub, (lo, hi) in zip(self.subs, self.bounds):
            sub.load(xa[lo:hi], [f[lo:hi] for f in feats])

    def configure_greedy(self, *args, **kw):
        seed = int(kw.pop("seed", 0))
        for i, sub in enumerate(self.subs):  # the sampling RNG is keyed on (seed, row): decorrelate the sub-batches
            sub.configure_greedy(*args, seed=seed + 0xA0761D6478BD642F * i, **kw)
        self.n_init = self.subs[0].n_init

    def set_initial_rows(self, rows: Tensor):
        for sub, (lo, hi) in zip(self.subs, self.bounds):
            sub.set_initial_rows(rows[lo:hi])

    def run_greedy(self, n_sample: int, check_every: int = 8) -> int:
        total = self.n_init - 1 + n_sample
        main = torch.cuda.current_stream(self.dev)
        for sub in self.subs:
            sub._ensure_graph()
        for st in self.streams:
            st.wait_stream(main)
        launches = 0
        for i in range(total):
            for sub, st in zip(self.subs, self.streams):
                with torch.cuda.stream(st):
                    sub.step()
            launches += 1
            sampled = i - (self.n_init - 1) + 1
            if sampled > 0 and sampled % check_every == 0 and i + 1 < total:
                done = True
                for sub, st in zip(self.subs, self.streams):
                    with torch.cuda.stream(st):
                        done = sub.all_done() and done
                if done:
                    break
        for st in self.streams:
            main.wait_stream(st)
        return launches

    def results(self, n_tokens: int):
        toks, lps, nsp = [], [], []
        for sub in self.subs:
            t, l, n = sub.results(n_tokens)
            toks += t
            lps += l
            nsp += n
        return toks, lps, nsp


_SESSION_CACHE = weakref.WeakKeyDictionary()  # decoder module -> its most recent sessions (most recent first)


def _session_cache_size() -> int:
    """Sessions kept per decoder (WF_SESSION_CACHE, default 3).  The long-form driver alternates between a few shapes
    (the batch shrinks as recordings end, fallback sub-batches, bucketed token capacities): each keeps its K/V arena,
    step buffers and captured CUDA graph."""
    return max(1, int(os.environ.get("WF_SESSION_CACHE", "3")))


def default_split(n_rows: int, greedy: bool) -> int:
    """Sub-batches stepping concurrently (SplitSession).  WF_DECODE_SPLIT overrides; beam search is never split."""
    if not greedy:
        return 1
    env = os.environ.get("WF_DECODE_SPLIT")
    if env:
        return max(1, int(env))
    # Measured on B200 (large-v2 AV, B=128, tools/split_probe.py): 1 -> 964 ms, 2 -> 1063 ms, 4 -> 1195 ms per decode
    # loop; with launch priorities 1018 / 1293 ms.  The attention CTAs fill every SM, so the other sub-batch's GEMM
    # CTAs (160-190 KB of shared memory each) only get in when an SM drains completely: no useful overlap yet.
    return 1


def get_session(dec, xa: Tensor, feats: Optional[Sequence[Tensor]], n_group: int, t_cap: int, n_split: int = 1):
    """One cached session per decoder: a repeated decode() with the same shapes reuses the ~(B x 370 MB) K/V arena,
    the step buffers and the captured CUDA graph instead of re-allocating and re-capturing them every call."""
    dt = _engine_dtype(xa.dtype)
    p = decoder_pack(dec, dt)
    n_split = max(1, min(n_split, xa.shape[0]))
    if len(p.blocks) > 0 and len(p.blocks[0].x_attn) > 0:  # refuse mismatched features before anything is sized by them
        feats = validate_features(feats, xa.shape[0], xa.device, len(p.blocks[0].x_attn))
    if os.environ.get("WF_NO_GRAPH", "0") == "1":
        n_split = 1  # per-kernel profiling mode: one eager stream
    n_feats = 0 if feats is None else len(feats)
    no_cache = os.environ.get("WF_NO_SESSION_CACHE", "0") == "1"
    if no_cache:  # profiling sessions (eager, instrumented) neither use nor disturb the cached product sessions
        if n_split > 1:
            return SplitSession(dec, xa, feats, n_group, t_cap, n_split)
        return DecodeSession(dec, xa, feats, n_group, t_cap)
    cached = _SESSION_CACHE.get(dec) or []
    for i, old in enumerate(cached):
        tx = tuple(f.shape[1] for f in feats) if (old.gated and feats is not None) else ()
        key = (dt, xa.device, xa.shape[0], n_group, t_cap, xa.shape[1], tx, id(p))
        if n_split > 1:
            key = key + (n_split,)
        same_path = (not isinstance(old, DecodeSession) or (
            old.latent == (old.fold is not None and latent_cross_enabled(old.R, old.G, p.n_head, p.d)) and
            old.latent_x == (old.latent and old.gated and os.environ.get("WF_LATENT_X", "1") != "0")))
        if old.shape_key() == key and (not old.gated or n_feats == len(old.Tx)) and same_path:
            cached.insert(0, cached.pop(i))
            old.load(xa, old._check_feats(feats))
            return old
    # sessions of a repacked decoder (weights changed) can never match again: drop them, then make room
    cached = [c for c in cached if c.p is p][: _session_cache_size() - 1]
    _SESSION_CACHE[dec] = cached
    if n_split > 1:
        sess = SplitSession(dec, xa, feats, n_group, t_cap, n_split)
    else:
        sess = DecodeSession(dec, xa, feats, n_group, t_cap)
    cached.insert(0, sess)
    return sess


def last_session(dec):
    """The most recently used cached session of a decoder (None when there is none) - test / tooling hook."""
    cached = _SESSION_CACHE.get(dec)
    return cached[0] if cached else None


def clear_sessions() -> None:
    _SESSION_CACHE.clear()


# ============================================================================ stand-alone sub-module calls
@torch.no_grad()
def standalone_layernorm(ln, x: Tensor) -> Tensor:
    nv.require_cuda(x)
    dt = _engine_dtype(x.dtype)
    x2 = _to_dtype(x, dt).reshape(-1, x.shape[-1]).contiguous()
    out = torch.empty_like(x2)
    with torch.cuda.device(x.device):
        nv.layernorm(x2, _f32(ln.weight), _f32(ln.bias), out, ln.eps)
    return out.view(x.shape).to(x.dtype)


@torch.no_grad()
def standalone_linear(lin, x: Tensor) -> Tensor:
    nv.require_cuda(x)
    dt = _engine_dtype(x.dtype)
    x2 = _to_dtype(x, dt).reshape(-1, x.shape[-1]).contiguous()
    with torch.cuda.device(x.device):
        w = _pack_w(lin.weight, dt)
        out = _empty(x2.shape[0], w.shape[0], dt, x.device)
        nv.linear(x2, w, out, bias=None if lin.bias is None else _f32(lin.bias))
    return out.view(*x.shape[:-1], w.shape[0]).to(x.dtype)


@torch.no_grad()
def standalone_mha(mha, x: Tensor, xa: Optional[Tensor], causal: bool) -> Tensor:
    """MultiHeadAttention.forward (reference model.py:71-91) without kv_cache."""
    nv.require_cuda(x)
    dt = _engine_dtype(x.dtype)
    with torch.cuda.device(x.device):
        mp = _pack_mha(mha, dt)
        B, t, d = x.shape
        src = x if xa is None else xa
        Ts = src.shape[1]
        x2 = _to_dtype(x, dt).reshape(B * t, d).contiguous()
        s2 = x2 if xa is None else _to_dtype(xa, dt).reshape(B * Ts, d).contiguous()
        q, kv = _empty(B * t, d, dt, x.device), _empty(B * Ts, 2 * d, dt, x.device)
        nv.linear(x2, mp.q_w, q, bias=mp.q_b)
        nv.linear(s2, mp.kv_w, kv, bias=mp.kv_b)
        att = _empty(B * t, d, dt, x.device)
        nv.attention(q, kv[:, :d], kv[:, d:], att, B, t, Ts, mha.n_head, causal=causal)
        out = _empty(B * t, d, dt, x.device)
        nv.linear(att, mp.o_w, out, bias=mp.o_b)
    return out.view(B, t, d).to(x.dtype)


@torch.no_grad()
def standalone_conv1d(conv, x: Tensor) -> Tensor:
    """Conv1d.forward (reference model.py:44-50) for the stem's geometry (kernel 3, padding 1, stride 1 | 2, no
    dilation / groups) as im2col + GEMM; x (batch, C_in, T) -> (batch, C_out, T_out), no activation."""
    nv.require_cuda(x)
    if (conv.kernel_size != (3,) or conv.padding != (1,) or conv.stride not in ((1,), (2,)) or conv.dilation != (1,)
            or conv.groups != 1 or conv.padding_mode != "zeros"):
        raise nv.WfError(f"Conv1d geometry {conv} is not the Whisper stem's (kernel 3, padding 1, stride 1|2)")
    if x.dim() != 3 or x.shape[1] != conv.in_channels:
        raise ValueError(f"expected (batch, {conv.in_channels}, T), got {tuple(x.shape)}")
    dt = _engine_dtype(x.dtype)
    with torch.cuda.device(x.device):
        xin = _to_dtype(x, dt).contiguous()
        B, C, T = xin.shape
        stride = conv.stride[0]
        T_out = (T - 1) // stride + 1
        cols = _empty(B * T_out, 3 * C, dt, x.device)
        nv.im2col_k3(xin, C * T, T, 1, B, C, T, stride, cols)
        out = _empty(B * T_out, conv.out_channels, dt, x.device)
        nv.linear(cols, _pack_w(conv.weight, dt), out, bias=None if conv.bias is None else _f32(conv.bias))
    return out.view(B, T_out, conv.out_channels).permute(0, 2, 1).to(x.dtype)


@torch.no_grad()
def standalone_gated_xattn(sub, x: Tensor, xt: Tensor) -> Tensor:
    """GatedXAttnSubBlock.forward (reference model.py:121-134): tanh(attn_gate) * MHA(LN(x), xt) - the delta only."""
    nv.require_cuda(x, xt)
    dt = _engine_dtype(x.dtype)
    B, t, d = x.shape
    if xt.dim() != 3 or xt.shape[0] != B or xt.shape[2] != d:
        raise ValueError(f"xt must be ({B}, T_x, {d}), got {tuple(xt.shape)}")
    with torch.cuda.device(x.device):
        mp, ln = _pack_mha(sub.attn, dt), _pack_ln(sub.attn_ln)
        Tx = xt.shape[1]
        x2 = _to_dtype(x, dt).reshape(B * t, d).contiguous()
        f2 = _to_dtype(xt, dt).reshape(B * Tx, d).contiguous()
        xn, q, att = torch.empty_like(x2), torch.empty_like(x2), torch.empty_like(x2)
        kv = _empty(B * Tx, 2 * d, dt, x.device)
        nv.layernorm(x2, ln.w, ln.b, xn)
        nv.linear(xn, mp.q_w, q, bias=mp.q_b)
        nv.linear(f2, mp.kv_w, kv, bias=mp.kv_b)
        nv.attention(q, kv[:, :d], kv[:, d:], att, B, t, Tx, sub.attn.n_head, causal=False)
        out = torch.empty_like(x2)
        nv.linear(att, mp.o_w, out, bias=mp.o_b, gate=_f32(sub.attn_gate))
    return out.view(B, t, d).to(x.dtype)


@torch.no_grad()
def standalone_block(blk, x: Tensor, xa: Optional[Tensor], causal: bool, xt_list: Optional[Sequence[Tensor]]) -> Tensor:
    """ResidualAttentionBlock.forward (reference model.py:201-215) without kv_cache.  ``xt_list`` holds feature tensors
    already brought to (batch, T_x, n_state) - at block level the reference receives them after TextDecoder's
    projection and positional add (model.py:313-326)."""
    nv.require_cuda(x, xa)
    dt = _engine_dtype(x.dtype)
    B, t, d = x.shape
    gated = blk.add_gated_x_attn != 0
    with torch.cuda.device(x.device):
        bp = _cached_pack(blk, dt, _pack_block)
        feats = None
        if gated:
            xt_list = validate_features(xt_list, B, x.device, len(bp.x_attn))
            for i, f in enumerate(xt_list):
                if f.shape[2] != d:
                    raise ValueError(f"feature tensor {i} must be (batch, T_x, {d}) at block level, got {tuple(f.shape)}")
            feats = [_to_dtype(f, dt).reshape(-1, d).contiguous() for f in xt_list]
        xa2, Ta = None, 0
        if blk.cross_attn is not None and xa is not None:
            if xa.dim() != 3 or xa.shape[0] != B or xa.shape[2] != d:
                raise ValueError(f"xa must be ({B}, T, {d}), got {tuple(xa.shape)}")
            xa2, Ta = _to_dtype(xa, dt).reshape(-1, d).contiguous(), xa.shape[1]
        x2 = _to_dtype(x, dt).reshape(B * t, d).clone()
        ws = _BlockScratch(B * t, d, dt, x.device)
        out = _block_forward(bp, ws, x2, B, t, blk.attn.n_head, xa2, Ta, feats, causal=causal)
    return out.view(B, t, d).to(x.dtype)


# ============================================================================ hook-style kv_cache compatibility path
def hooks_in_use(dec, kv_cache: Optional[dict]) -> bool:
    """True when a caller drives the decoder the reference's hook way (Whisper.install_kv_cache_hooks, reference
    model.py:394-425): a non-empty cache dictionary, or forward hooks sitting on the key / value projections."""
    if kv_cache is None:
        return False
    if len(kv_cache) > 0:
        return True
    blk = dec.blocks[0] if len(dec.blocks) else None
    return blk is not None and len(blk.attn.key._forward_hooks) > 0


def _mha_with_cache(m, mp: _MhaPack, xin: Tensor, x_res: Optional[Tensor], src: Optional[Tensor], causal: bool,
                    kv_cache: dict, gate: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """MultiHeadAttention.forward with the reference's kv_cache protocol (model.py:71-91): the key / value projections
    are MODULE calls, so installed forward hooks see (and replace) their outputs; everything else is engine calls.
    xin [B, t, d] (already normalised), x_res [B * t, d] residual rows updated in place (or `out` += when given)."""
    B, t, d = xin.shape
    dt, dev = xin.dtype, xin.device
    q = _empty(B * t, d, dt, dev)
    nv.linear(xin.reshape(B * t, d), mp.q_w, q, bias=mp.q_b)
    if src is None or m.key not in kv_cache:
        s3 = xin if src is None else src
        k, v = m.key(s3), m.value(s3)       # hooks (if installed) return the concatenated cache tensors
    else:
        k, v = kv_cache[m.key], kv_cache[m.value]
    Tk = k.shape[1]
    k2, v2 = k.reshape(B * Tk, d).contiguous(), v.reshape(B * Tk, d).contiguous()
    att = _empty(B * t, d, dt, dev)
    # the reference adds mask[:t, :t] to a [t, Tk] score matrix: causal when t == Tk, a no-op 0 for one new token
    nv.attention(q, k2, v2, att, B, t, Tk, m.n_head, causal=causal and t == Tk and t > 1)
    if x_res is None and out is None:  # stand-alone module call: no residual stream
        return nv.linear(att, mp.o_w, _empty(B * t, d, dt, dev), bias=mp.o_b, gate=gate)
    dst = x_res if out is None else out
    nv.linear(att, mp.o_w, dst, bias=mp.o_b, residual=dst, gate=gate)
    return dst


@torch.no_grad()
def standalone_mha_cached(mha, x: Tensor, xa: Optional[Tensor], causal: bool, kv_cache: dict) -> Tensor:
    """MultiHeadAttention.forward with a reference-style kv_cache dictionary (model.py:71-91)."""
    nv.require_cuda(x, xa)
    if x.dtype not in _ENGINE_DTYPES:
        raise nv.WfError("the kv_cache hook path runs in float32 or bfloat16 activations")
    with torch.cuda.device(x.device):
        mp = _pack_mha(mha, x.dtype)
        out = _mha_with_cache(mha, mp, x.contiguous(), None, None if xa is None else xa.contiguous(), causal, kv_cache)
    return out.view(x.shape)


@torch.no_grad()
def decoder_forward_hooked(dec, tokens: Tensor, xa: Tensor, kv_cache: dict, xt_list: Optional[Sequence[Tensor]]):
    """TextDecoder.forward driven through reference-style kv_cache hooks (model.py:292-340 with a cache dictionary):
    positions start at the cached length, key / value projections go through the hooked modules.  A compatibility
    path for callers that manage the cache themselves; whisper.decode() uses the engine's own sessions instead."""
    nv.require_cuda(tokens, xa)
    if xa.dtype not in _ENGINE_DTYPES:
        raise nv.WfError("the kv_cache hook path runs in float32 or bfloat16 activations")
    dt = xa.dtype
    with torch.cuda.device(xa.device):
        p = decoder_pack(dec, dt)
        gated = len(p.blocks) > 0 and len(p.blocks[0].x_attn) > 0
        B, t = tokens.shape
        d, dev = p.d, xa.device
        offset = next(iter(kv_cache.values())).shape[1] if kv_cache else 0
        if offset + t > p.n_ctx:
            raise RuntimeError(f"positions {offset}..{offset + t} exceed n_text_ctx {p.n_ctx}")
        feats3 = None
        if gated:
            xt_list = validate_features(xt_list, B, dev, len(p.blocks[0].x_attn))
            feats3 = []
            for xt in xt_list:
                Tx = xt.shape[1]
                if offset + Tx > p.n_ctx:
                    raise RuntimeError(f"feature positions {offset}..{offset + Tx} exceed n_text_ctx {p.n_ctx}")
                if xt.shape[2] != d:
                    a = _to_dtype(xt.reshape(B * Tx, -1).float() if xt.dtype == torch.float16 else
                                  xt.reshape(B * Tx, -1), dt).contiguous()
                    f = nv.linear(a, p.xt_w, _empty(B * Tx, d, dt, dev), bias=p.xt_b,
                                  residual=p.pos_emb_t[offset:offset + Tx], res_row_mod=Tx)
                else:
                    f = nv.add_rowmod(xt.reshape(B * Tx, d).contiguous(), p.pos_emb[offset:].contiguous(),
                                      _empty(B * Tx, d, dt, dev), Tx)
                feats3.append(f.view(B, Tx, d))
        tok32 = tokens.to(torch.int32).contiguous()
        x = _empty(B * t, d, dt, dev)
        nv.embed(tok32, t, None, 0, p.tok_emb, p.pos_emb[offset:].contiguous(), x, n_pos=t)
        xn, hbuf = _empty(B * t, d, dt, dev), _empty(B * t, 4 * d, dt, dev)
        xa3 = xa.contiguous()
        for blk, bp in zip(dec.blocks, p.blocks):
            if gated:
                acc = x if len(feats3) <= 1 else x.clone()
                for i, f in enumerate(feats3):
                    nv.layernorm(x, bp.x_ln[i].w, bp.x_ln[i].b, xn)
                    _mha_with_cache(blk.gated_x_attn_layers[i].attn, bp.x_attn[i], xn.view(B, t, d), x, f, False,
                                    kv_cache, gate=bp.x_gate[i], out=acc)
                x = acc
                _mlp_inplace(x, xn, hbuf, bp.ff_ln, bp.ff, gate=bp.ff_gate)
            nv.layernorm(x, bp.attn_ln.w, bp.attn_ln.b, xn)
            _mha_with_cache(blk.attn, bp.attn, xn.view(B, t, d), x, None, True, kv_cache)
            nv.layernorm(x, bp.cross_ln.w, bp.cross_ln.b, xn)
            _mha_with_cache(blk.cross_attn, bp.cross, xn.view(B, t, d), x, xa3, False, kv_cache)
            _mlp_inplace(x, xn, hbuf, bp.mlp_ln, bp.mlp)
        nv.layernorm(x, p.ln.w, p.ln.b, xn)
        logits = torch.empty((B * t, p.n_vocab), dtype=torch.float32, device=dev)
        nv.linear(xn, p.tok_emb_t, logits)
    return logits.view(B, t, p.n_vocab)
