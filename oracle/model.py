"""Oracle: Whisper-Flamingo network (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional torch-CPU restatement of reference ``whisper/model.py``.  It owns no
parameters: every function takes ``sd``, a ``{state_dict name: tensor}`` mapping
with the reference's own key names (model.py:342-429), so the very
same weights can be fed to the reference, to this oracle and to the CUDA engine.

dtype policy (model.py:30-50): parameters stay fp32 ("master"),
``Linear``/``Conv1d`` cast them to the activation dtype at every call,
``LayerNorm`` and softmax compute in fp32 and cast back.  ``act_dtype`` selects
the activation dtype (fp32 = the token-exact path; bf16 = the reference's
half-precision semantics used to calibrate tolerances).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class Dims:
    """model.py:16-27 (ModelDimensions)."""
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


def _linear(sd, prefix: str, x: Tensor) -> Tensor:
    """model.py:35-41 - weight/bias cast to x.dtype on every call."""
    w = sd[prefix + ".weight"].to(x.dtype)
    b = sd.get(prefix + ".bias")
    return F.linear(x, w, None if b is None else b.to(x.dtype))


def _layer_norm(sd, prefix: str, x: Tensor) -> Tensor:
    """model.py:30-32 - fp32 LayerNorm (eps 1e-5), cast back."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5).type(x.dtype)


def sinusoids(length: int, channels: int, max_timescale: int = 10000) -> Tensor:
    """model.py:53-59."""
    assert channels % 2 == 0
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    st = torch.arange(length)[:, np.newaxis] * inv[np.newaxis, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def mha(sd, prefix: str, n_head: int, x: Tensor, xa: Optional[Tensor] = None,
        mask: Optional[Tensor] = None) -> Tensor:
    """model.py:62-108 (no kv_cache: the reference disables it).

    q = Wq x + b ; k = Wk src (no bias) ; v = Wv src + b ;
    softmax_fp32((q s)(k s)^T + mask) v with s = d_head^-0.25 ; out proj.
    """
    src = x if xa is None else xa
    q = _linear(sd, prefix + ".query", x)
    k = _linear(sd, prefix + ".key", src)
    v = _linear(sd, prefix + ".value", src)
    n_batch, n_ctx, n_state = q.shape
    scale = (n_state // n_head) ** -0.25
    q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3) * scale
    k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 3, 1) * scale
    v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    qk = q @ k
    if mask is not None:
        qk = qk + mask[:n_ctx, :n_ctx]
    w = F.softmax(qk.float(), dim=-1).to(q.dtype)
    wv = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
    return _linear(sd, prefix + ".out", wv)


def _mlp(sd, prefix: str, x: Tensor) -> Tensor:
    """nn.Sequential(Linear, GELU(exact erf), Linear) - model.py:149-152, 165-168."""
    return _linear(sd, prefix + ".2", F.gelu(_linear(sd, prefix + ".0", x)))


def _gate(sd, name: str, like: Tensor) -> Tensor:
    # The reference multiplies by a shape-[1] fp32 parameter; in half precision the
    # upstream recipe casts the gate to the activation dtype first.
    return sd[name].to(like.dtype).tanh()


def residual_block(sd, prefix: str, n_head: int, x: Tensor, xa: Optional[Tensor] = None,
                   mask: Optional[Tensor] = None, xt_list: Optional[List[Tensor]] = None,
                   cross: bool = False, gated: bool = False) -> Tensor:
    """model.py:136-215.  Decoder order: gated x-attn (+gated FF) -> self -> cross -> MLP."""
    if gated:
        # apply_gated_x_attn_multi, model.py:171-199 (+ GatedXAttnSubBlock :110-134)
        n_langs = sum(1 for k in sd if k.startswith(prefix + ".gated_x_attn_layers.") and k.endswith(".attn_gate"))
        if len(xt_list) > n_langs:
            raise ValueError(f"Got {len(xt_list)} translations but only support up to {n_langs}")
        x_origin = x
        total = 0
        for i, xt in enumerate(xt_list):
            p = f"{prefix}.gated_x_attn_layers.{i}"
            x_ln = _layer_norm(sd, p + ".attn_ln", x_origin)
            total = total + mha(sd, p + ".attn", n_head, x_ln, xt) * _gate(sd, p + ".attn_gate", x_origin)
        x = x_origin + total
        x = x + _mlp(sd, prefix + ".ff", _layer_norm(sd, prefix + ".ff_ln", x)) * _gate(sd, prefix + ".ff_gate", x)
    x = x + mha(sd, prefix + ".attn", n_head, _layer_norm(sd, prefix + ".attn_ln", x), mask=mask)
    if cross:
        x = x + mha(sd, prefix + ".cross_attn", n_head, _layer_norm(sd, prefix + ".cross_attn_ln", x), xa)
    x = x + _mlp(sd, prefix + ".mlp", _layer_norm(sd, prefix + ".mlp_ln", x))
    return x


def encoder_stem(sd, mel: Tensor) -> Tensor:
    """model.py:239-250: gelu(conv1) -> gelu(conv2, stride 2) -> [B,T,d] -> crop -> + sinusoids."""
    def conv(prefix, x, stride):
        return F.conv1d(x, sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype),
                        stride=stride, padding=1)
    x = F.gelu(conv("encoder.conv1", mel, 1))
    x = F.gelu(conv("encoder.conv2", x, 2))
    x = x.permute(0, 2, 1)
    if x.shape[1] > 1500:
        x = x[:, :1500, :]
    pos = sd["encoder.positional_embedding"]
    return (x + pos[: x.shape[1]]).to(x.dtype)


def encoder_forward(sd, dims: Dims, mel: Tensor, act_dtype=torch.float32,
                    return_layers: bool = False):
    """AudioEncoder.forward, model.py:234-258.  mel [B, n_mels, T<=3000] -> [B, T/2, d]."""
    x = encoder_stem(sd, mel.to(act_dtype))
    outs = [x]
    for i in range(dims.n_audio_layer):
        x = residual_block(sd, f"encoder.blocks.{i}", dims.n_audio_head, x)
        outs.append(x)
    x = _layer_norm(sd, "encoder.ln_post", x)
    return (x, outs) if return_layers else x


def has_gated_x_attn(sd) -> bool:
    return any(".gated_x_attn_layers." in k for k in sd)


def decoder_forward(sd, dims: Dims, tokens: Tensor, xa: Tensor,
                    xt_list: Optional[List[Tensor]] = None, return_layers: bool = False):
    """TextDecoder.forward, model.py:292-340, kv_cache empty (offset 0).

    tokens [B,t] int64, xa [B,1500,d], xt_list [[B,T_x,bert_dim]] -> logits fp32 [B,t,n_vocab].
    """
    pos = sd["decoder.positional_embedding"]
    x = F.embedding(tokens, sd["decoder.token_embedding.weight"]) + pos[: tokens.shape[-1]]
    x = x.to(xa.dtype)
    gated = has_gated_x_attn(sd)
    processed = None
    if xt_list is not None:
        processed = []
        for xt in xt_list:
            if xt.shape[-1] != x.shape[-1]:
                # plain nn.Linear (NOT the casting Linear), model.py:286-290, 318-319
                xt = F.linear(xt, sd["decoder.xt_projection.weight"], sd["decoder.xt_projection.bias"])
            xt = xt + pos[: xt.shape[1]]  # learned *text* pos-emb, model.py:322
            processed.append(xt.to(xa.dtype))
    n_ctx = dims.n_text_ctx
    mask = torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1)
    outs = [x]
    for i in range(dims.n_text_layer):
        x = residual_block(sd, f"decoder.blocks.{i}", dims.n_text_head, x, xa, mask=mask,
                           xt_list=processed, cross=True, gated=gated)
        outs.append(x)
    x = _layer_norm(sd, "decoder.ln", x)
    logits = (x @ torch.transpose(sd["decoder.token_embedding.weight"].to(x.dtype), 0, 1)).float()
    return (logits, outs) if return_layers else logits


def cast_state_dict_fp32(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    return {k: v.detach().to("cpu", torch.float32) if v.is_floating_point() else v.detach().cpu()
            for k, v in sd.items()}
