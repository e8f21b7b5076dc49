"""Model classes of the drop-in ``whisper`` package.

Public surface and ``state_dict`` key names are those of reference ``whisper/model.py``
(``ModelDimensions`` :16-27, ``LayerNorm/Linear/Conv1d`` :30-50, ``MultiHeadAttention`` :62-108,
``GatedXAttnSubBlock`` :110-134, ``ResidualAttentionBlock`` :136-215, ``AudioEncoder`` :217-258,
``TextDecoder`` :260-340, ``Whisper`` :342-429), so checkpoints and the fork's training scripts that
poke at sub-modules keep working.  The modules only *hold* the fp32 master parameters: every
``forward`` hands the tensors to the CUDA engine (``_engine.py`` -> ``libwf.so``).  There is no
PyTorch implementation of the math here and no CPU path.

Activation dtype selects the engine: float32 -> token-exact CUDA-core engine, bfloat16 -> tcgen05
tensor-core engine (float16 inputs are computed on the bf16 engine and cast back).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _engine
from .decoding import decode as decode_function
from .decoding import detect_language as detect_language_function
from .transcribe import transcribe as transcribe_function


@dataclass
class ModelDimensions:
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


class LayerNorm(nn.LayerNorm):
    """Parameter holder; fp32 statistics are computed by ``wf_layernorm``."""

    def forward(self, x: Tensor) -> Tensor:
        return _engine.standalone_layernorm(self, x)


class Linear(nn.Linear):
    """Parameter holder; executed by ``wf_linear``."""

    def forward(self, x: Tensor) -> Tensor:
        return _engine.standalone_linear(self, x)


class Conv1d(nn.Conv1d):
    """Parameter holder; the stem geometry (kernel 3, padding 1, stride 1 | 2) runs as im2col + ``wf_linear``."""

    def forward(self, x: Tensor) -> Tensor:
        return _engine.standalone_conv1d(self, x)


def sinusoids(length: int, channels: int, max_timescale: int = 10000) -> Tensor:
    """Sinusoidal positional embedding [length, channels] = cat(sin, cos) (reference model.py:53-59)."""
    assert channels % 2 == 0
    half = channels // 2
    log_inc = np.log(max_timescale) / (half - 1)
    inv_timescales = torch.exp(-log_inc * torch.arange(half))
    angles = torch.arange(length)[:, None] * inv_timescales[None, :]
    return torch.cat([angles.sin(), angles.cos()], dim=1)


class MultiHeadAttention(nn.Module):
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None):
        """Returns ``(out, None)``: the fused attention never materialises the fp32 score tensor the
        reference returns as its second value (only the broken word-timestamp path used it)."""
        if kv_cache is not None and (len(kv_cache) > 0 or len(self.key._forward_hooks) > 0):
            return _engine.standalone_mha_cached(self, x, xa, mask is not None, kv_cache), None
        return _engine.standalone_mha(self, x, xa, causal=mask is not None), None


class GatedXAttnSubBlock(nn.Module):
    """tanh(attn_gate) * MHA(LN(x), xt) - the delta only, no residual (reference model.py:110-134)."""

    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.attn_gate = nn.Parameter(torch.tensor([0.0]))

    def forward(self, x: Tensor, xt: Tensor) -> Tensor:
        return _engine.standalone_gated_xattn(self, x, xt)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_state: int, n_head: int, cross_attention: bool = False, add_adapter: bool = False,
                 adapter_dim: int = 256, add_gated_x_attn: int = 0, num_langs: int = 0):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.cross_attn = MultiHeadAttention(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state) if cross_attention else None
        n_mlp = n_state * 4
        self.mlp = nn.Sequential(Linear(n_state, n_mlp), nn.GELU(), Linear(n_mlp, n_state))
        self.mlp_ln = LayerNorm(n_state)
        self.add_gated_x_attn = add_gated_x_attn
        self.num_langs = num_langs
        if self.add_gated_x_attn != 0:
            self.gated_x_attn_layers = nn.ModuleList(GatedXAttnSubBlock(n_state, n_head) for _ in range(num_langs))
            self.ff_ln = LayerNorm(n_state)
            self.ff = nn.Sequential(Linear(n_state, n_mlp), nn.GELU(), Linear(n_mlp, n_state))
            self.ff_gate = nn.Parameter(torch.tensor([0.0]))

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None, xt_list: Optional[List[Tensor]] = None):
        """One block over full sequences (reference model.py:201-215).  ``mask`` only selects causal self-attention
        (the reference's mask is always the causal upper-triangular one, model.py:281); hook-style ``kv_cache``
        dictionaries are not supported (the engine owns the KV cache of whisper.decode)."""
        if kv_cache:
            raise RuntimeError("hook-based kv_cache dictionaries are not supported at block level")
        return _engine.standalone_block(self, x, xa, mask is not None, xt_list)


class AudioEncoder(nn.Module):
    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int,
                 dropout_rate: float, add_adapter: bool, adapter_dim: int):
        super().__init__()
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            ResidualAttentionBlock(n_state, n_head, False, add_adapter, adapter_dim, add_gated_x_attn=0)
            for _ in range(n_layer))
        self.ln_post = LayerNorm(n_state)
        self.dropout_rate = dropout_rate
        self.dropout = nn.Dropout(dropout_rate)  # held for state/API parity; never applied (as in the reference)
        self.n_head = n_head

    def forward(self, x: Tensor, track_norm: bool = False, padding_mask=None):
        """x: (batch, n_mels, n_frames <= 3000) mel -> (batch, n_frames // 2, n_state)."""
        return _engine.encoder_forward(self, x, track_norm=track_norm)


class TextDecoder(nn.Module):
    def __init__(self, n_vocab: int, n_ctx: int, n_state: int, n_head: int, n_layer: int, dropout_rate: float,
                 add_gated_x_attn: int, bert_hidden_size: int, num_langs: int):
        super().__init__()
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        self.positional_embedding = nn.Parameter(torch.empty(n_ctx, n_state))
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            ResidualAttentionBlock(n_state, n_head, cross_attention=True, add_gated_x_attn=add_gated_x_attn,
                                   num_langs=num_langs) for _ in range(n_layer))
        self.ln = LayerNorm(n_state)
        mask = torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1)
        self.register_buffer("mask", mask, persistent=False)
        self.dropout_rate = dropout_rate
        self.dropout = nn.Dropout(dropout_rate)
        # features whose width differs from n_state go through a plain projection first
        self.xt_projection = nn.Linear(bert_hidden_size, n_state) if bert_hidden_size != n_state else nn.Identity()
        self.n_head = n_head
        self.n_ctx = n_ctx

    def forward(self, x: Tensor, xa: Tensor, kv_cache: Optional[dict] = None,
                xt_list: Optional[List[Tensor]] = None):
        """x: (batch, t <= n_ctx) token ids; xa: (batch, n_audio_ctx, n_state) encoder output;
        xt_list: up to ``num_langs`` feature tensors (batch, T_x, bert_dim | n_state) for the gated
        cross-attention.  Returns fp32 logits (batch, t, n_vocab).  ``kv_cache``: a dictionary filled by
        ``Whisper.install_kv_cache_hooks`` switches to the reference's hook protocol (positions start at the cached
        length); ``whisper.decode`` does not need it - its caching lives inside the engine's decode sessions."""
        if _engine.hooks_in_use(self, kv_cache):
            return _engine.decoder_forward_hooked(self, x, xa, kv_cache, xt_list)
        return _engine.decoder_forward(self, x, xa, xt_list)


class Whisper(nn.Module):
    def __init__(self, dims: ModelDimensions, dropout_rate: float = 0.0, add_adapter: bool = False,
                 adapter_dim: int = 256, add_gated_x_attn: int = 0, bert_dim: int = 768, num_langs: int = 0):
        super().__init__()
        self.dims = dims
        self.encoder = AudioEncoder(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head,
                                    dims.n_audio_layer, dropout_rate, add_adapter, adapter_dim)
        self.decoder = TextDecoder(dims.n_vocab, dims.n_text_ctx, dims.n_text_state, dims.n_text_head,
                                   dims.n_text_layer, dropout_rate, add_gated_x_attn, bert_dim, num_langs)

    def embed_audio(self, mel: Tensor):
        return self.encoder(mel)

    def logits(self, tokens: Tensor, audio_features: Tensor):
        return self.decoder(tokens, audio_features)

    def forward(self, mel: Tensor, tokens: Tensor) -> Dict[str, Tensor]:
        return self.decoder(tokens, self.encoder(mel))

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def is_multilingual(self):
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self):
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)

    def install_kv_cache_hooks(self, cache: Optional[dict] = None):
        """Reference-style KV cache (model.py:394-425): returns ``(cache, hooks)``; forward hooks on every key / value
        projection of the decoder keep their outputs in ``cache`` - stored as-is the first time (and for anything
        longer than n_text_ctx, i.e. cross-attention over the audio), appended along time afterwards - and hand the
        cached tensor back as the projection's output.  ``whisper.decode`` does not use this (the engine owns its KV
        cache); it exists for callers that step ``model.decoder(tokens, xa, kv_cache=cache)`` themselves.
        As in the reference the hooks attach to EVERY MultiHeadAttention of the decoder, so they are only meaningful
        for audio-only models."""
        cache = dict(cache) if cache is not None else {}
        hooks = []
        n_ctx = self.dims.n_text_ctx

        def keep(module, _inputs, output):
            if module in cache and output.shape[1] <= n_ctx:
                cache[module] = torch.cat([cache[module], output], dim=1).detach()
            else:
                cache[module] = output
            return cache[module]

        for layer in self.decoder.modules():
            if isinstance(layer, MultiHeadAttention):
                hooks.append(layer.key.register_forward_hook(keep))
                hooks.append(layer.value.register_forward_hook(keep))
        return cache, hooks

    detect_language = detect_language_function
    transcribe = transcribe_function
    decode = decode_function
