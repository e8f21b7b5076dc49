"""B200-native drop-in for the ``whisper`` package of jerryyang1231/whisper-flamingo.

Import surface of reference ``whisper/__init__.py:11-15`` for the audio(-visual) inference hot
path: ``log_mel_spectrogram`` / ``pad_or_trim``, ``ModelDimensions`` / ``Whisper``,
``DecodingOptions`` / ``DecodingResult`` / ``decode`` / ``detect_language``, ``load_model``.
All numeric work runs in ``libwf.so`` (hand-written sm_100a CUDA, see ``include/wf.h``).
"""
from __future__ import annotations

import io
import os
from typing import List, Optional, Union

import torch

from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult, decode, detect_language
from .model import ModelDimensions, Whisper
from .transcribe import transcribe
from .version import __version__

# official checkpoint names (weights must already be on disk: this build never downloads)
_MODEL_NAMES = ("tiny.en", "tiny", "base.en", "base", "small.en", "small", "medium.en", "medium",
                "large-v1", "large-v2", "large-v3", "large")


def available_models() -> List[str]:
    return list(_MODEL_NAMES)


def load_model(name: str, device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, dropout_rate: float = 0.0, add_adapter: bool = False,
               adapter_dim: int = 256, add_gated_x_attn: int = 0, bert_dim: int = 768, num_langs: int = 0
               ) -> Whisper:
    """Load a Whisper(-Flamingo) checkpoint (reference ``whisper/__init__.py:99-164``).

    ``name`` is a path to a checkpoint ``{"dims": ..., "model_state_dict": ...}`` or an official
    model name resolved inside ``download_root`` (default ``~/.cache/whisper``) as ``<name>.pt``;
    nothing is downloaded.  Keys absent from the checkpoint (e.g. the gated x-attn layers when an
    audio-only checkpoint is extended) keep their initial values (``strict=False``), as in the
    reference.
    """
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if download_root is None:
        default = os.path.join(os.path.expanduser("~"), ".cache")
        download_root = os.path.join(os.getenv("XDG_CACHE_HOME", default), "whisper")
    if os.path.isfile(name):
        path = name
    elif name in _MODEL_NAMES:
        path = os.path.join(download_root, f"{name}.pt")
        if not os.path.isfile(path):
            raise RuntimeError(f"Model {name} not found at {path}; this build does not download checkpoints")
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if in_memory:
        with open(path, "rb") as fh:
            checkpoint = torch.load(io.BytesIO(fh.read()), map_location=device)
    else:
        checkpoint = torch.load(path, map_location=device)
    dims = ModelDimensions(**checkpoint["dims"])
    model = Whisper(dims, dropout_rate, add_adapter, adapter_dim, add_gated_x_attn, bert_dim, num_langs)
    model.load_state_dict(checkpoint["model_state_dict"], strict=False)
    return model.to(device)
