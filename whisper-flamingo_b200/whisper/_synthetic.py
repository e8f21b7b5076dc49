"""Deterministic random initialisation of a Whisper(-Flamingo) model (no checkpoints are reachable).

Every parameter gets its own ``torch.Generator`` seeded from ``seed`` and the CRC32 of its
state_dict name, so the reference model, the oracle and this package receive identical values
regardless of construction order (weights std 0.02, gates 0.5,
``decoder.positional_embedding`` std 0.01; 1-D parameters are also randomised here so that
biases and LayerNorm affine terms are exercised by the parity tests).
"""
from __future__ import annotations

import zlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


def synthetic_value(name: str, shape, seed: int = 0) -> torch.Tensor:
    g = _gen(seed, name)
    if name.endswith("attn_gate") or name.endswith("ff_gate"):
        return torch.full(shape, 0.5)
    if name == "decoder.positional_embedding":
        return torch.randn(shape, generator=g) * 0.01
    if len(shape) >= 2:
        return torch.randn(shape, generator=g) * 0.02
    if name.endswith(".weight") and ("_ln" in name or ".ln" in name or name.startswith("ln")):
        return 1.0 + torch.randn(shape, generator=g) * 0.02  # LayerNorm gain
    return torch.randn(shape, generator=g) * 0.02  # biases


@torch.no_grad()
def init_synthetic_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    for name, p in model.named_parameters():
        p.copy_(synthetic_value(name, tuple(p.shape), seed).to(p.dtype))
    return model


def synthetic_pcm(n_clips: int, n_samples: int = 480000, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n_clips, n_samples, generator=g) * 0.1


def synthetic_features(n_clips: int, n_frames: int = 750, dim: int = 1024, seed: int = 4321) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n_clips, n_frames, dim, generator=g)


@torch.no_grad()
def init_synthetic_fast_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Same distributions as ``init_synthetic_`` but drawn on the parameters' own device (benchmarks of the
    big models: values need not be reproducible across devices, only well-conditioned)."""
    for name, p in model.named_parameters():
        g = torch.Generator(device=p.device)
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith("attn_gate") or name.endswith("ff_gate"):
            p.fill_(0.5)
        elif name == "decoder.positional_embedding":
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.01)
        elif p.dim() < 2 and name.endswith(".weight") and ("_ln" in name or ".ln" in name):
            p.copy_(1.0 + torch.randn(p.shape, generator=g, device=p.device) * 0.02)
        else:
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.02)
    return model
