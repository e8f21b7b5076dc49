"""B200-native drop-in for the ``whisper`` package of jerryyang1231/whisper-flamingo.

Import surface of reference ``whisper/__init__.py:11-15`` for the audio(-visual) inference hot
path: ``log_mel_spectrogram`` / ``pad_or_trim``, ``ModelDimensions`` / ``Whisper``,
``DecodingOptions`` / ``DecodingResult`` / ``decode`` / ``detect_language``, ``load_model``.
All numeric work runs in ``libwf.so`` (hand-written sm_100a CUDA, see ``include/wf.h``).
"""
from __future__ import annotations

import io
import re
import os
from typing import Dict, List, Optional, Union

import torch

from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult, decode, detect_language
from .model import ModelDimensions, Whisper
from .transcribe import transcribe, transcribe_batch
from .version import __version__

# official checkpoint names (weights must already be on disk: this build never downloads)
_MODEL_NAMES = ("tiny.en", "tiny", "base.en", "base", "small.en", "small", "medium.en", "medium",
                "large-v1", "large-v2", "large-v3", "large")


def available_models() -> List[str]:
    return list(_MODEL_NAMES)


_UPSTREAM_XATTN = re.compile(r"^(decoder\.blocks\.\d+\.)(gated_x_attn_ln\.|gated_x_attn\.|attn_gate$)(.*)$")


def remap_checkpoint_keys(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Checkpoint key names -> this package's (= the fork's) ``state_dict`` names.

    * Lightning checkpoints of the fork's training scripts prefix every key with ``model.``
      (reference ``whisper-flamingo_kloka_crawled.py:172-183`` strips it by hand);
    * upstream Whisper-Flamingo AV checkpoints keep the gated cross-attention directly in the block
      (``blocks.N.gated_x_attn.*``, ``gated_x_attn_ln.*``, ``attn_gate``; key list in
      ``notebooks/whisper_flamingo_demo.ipynb:7330``) where the fork has
      ``blocks.N.gated_x_attn_layers.0.{attn.*, attn_ln.*, attn_gate}``.
    Keys that already use the fork's names are passed through unchanged."""
    out = {}
    for key, value in state_dict.items():
        if key.startswith("model."):
            key = key[len("model."):]
        m = _UPSTREAM_XATTN.match(key)
        if m:
            part = {"gated_x_attn.": "attn.", "gated_x_attn_ln.": "attn_ln.", "attn_gate": "attn_gate"}[m.group(2)]
            key = f"{m.group(1)}gated_x_attn_layers.0.{part}{m.group(3)}"
        out[key] = value
    return out


def infer_dims(state_dict: Dict[str, torch.Tensor]) -> ModelDimensions:
    """ModelDimensions from tensor shapes, for checkpoints that do not carry a ``dims`` entry (Lightning)."""
    def count(prefix):
        return 1 + max(int(k[len(prefix):].split(".")[0]) for k in state_dict if k.startswith(prefix))
    d_a = state_dict["encoder.conv1.weight"].shape[0]
    d_t = state_dict["decoder.token_embedding.weight"].shape[1]
    return ModelDimensions(
        n_mels=state_dict["encoder.conv1.weight"].shape[1], n_audio_ctx=state_dict["encoder.positional_embedding"].shape[0],
        n_audio_state=d_a, n_audio_head=d_a // 64, n_audio_layer=count("encoder.blocks."),
        n_vocab=state_dict["decoder.token_embedding.weight"].shape[0],
        n_text_ctx=state_dict["decoder.positional_embedding"].shape[0], n_text_state=d_t, n_text_head=d_t // 64,
        n_text_layer=count("decoder.blocks."))


def load_model(name: str, device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, dropout_rate: float = 0.0, add_adapter: bool = False,
               adapter_dim: int = 256, add_gated_x_attn: int = 0, bert_dim: int = 768, num_langs: int = 0
               ) -> Whisper:
    """Load a Whisper(-Flamingo) checkpoint (reference ``whisper/__init__.py:99-164``).

    ``name`` is a path to a checkpoint ``{"dims": ..., "model_state_dict": ...}`` or an official
    model name resolved inside ``download_root`` (default ``~/.cache/whisper``) as ``<name>.pt``;
    nothing is downloaded.  Lightning checkpoints (``state_dict`` with a ``model.`` prefix, no ``dims``) and upstream
    Whisper-Flamingo key names are accepted too (``remap_checkpoint_keys``).  Keys absent from the checkpoint (e.g. the gated x-attn layers when an
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
    # {"dims", "model_state_dict"} (OpenAI / fork format) or a Lightning checkpoint {"state_dict": {"model.*": ...}}
    weights = checkpoint.get("model_state_dict", checkpoint.get("state_dict"))
    if weights is None:
        raise RuntimeError(f"{path}: neither 'model_state_dict' nor 'state_dict' in the checkpoint")
    weights = remap_checkpoint_keys(weights)
    dims = ModelDimensions(**checkpoint["dims"]) if "dims" in checkpoint else infer_dims(weights)
    model = Whisper(dims, dropout_rate, add_adapter, adapter_dim, add_gated_x_attn, bert_dim, num_langs)
    model.load_state_dict(weights, strict=False)
    return model.to(device)
