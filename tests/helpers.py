"""Shared helpers for the test-suite (model construction, specs, golden loading)."""
import json
import os

import numpy as np
import torch

from oracle import decode as odec
from oracle import model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE_IDX = np.arange(0, 80 * 3000, 977)

TINY = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4, n_vocab=51865,
            n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)


def load_decode_golden():
    with open(os.path.join(GOLDEN, "decode_tiny.json")) as fh:
        return json.load(fh)


def spec_from_json(d) -> odec.DecodeSpec:
    d = dict(d)
    for k in ("initial_tokens", "suppress_tokens", "blank_tokens"):
        d[k] = tuple(d[k])
    return odec.DecodeSpec(**d)


def build_model(gated: bool, device="cpu", dims=None, bert_dim=1024, num_langs=1, seed=0):
    import whisper
    from whisper._synthetic import init_synthetic_
    dims = whisper.ModelDimensions(**(dims or TINY))
    if gated:
        model = whisper.Whisper(dims, 0.0, False, 256, 1, bert_dim, num_langs)
    else:
        model = whisper.Whisper(dims, 0.0, False, 256, 0, 768, 0)
    init_synthetic_(model, seed=seed)
    return model.eval().to(device)


def oracle_sd(model):
    return om.cast_state_dict_fp32(model.state_dict())


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())
