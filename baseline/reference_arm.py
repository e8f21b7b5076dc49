"""Drive the UNMODIFIED reference implementation for timing (bench.py `--impl reference`, `cpu_baseline`, and the
GPU-eager comparator).  NOT product code; nothing under whisper-flamingo_b200/ imports this.

The reference package is looked for at baseline/_ref/whisper (staged copy, see stage_reference.py) and then at
/root/reference/whisper, and imported under the alias `ref_whisper` so that it never shadows the drop-in `whisper`
package.  What is timed is the reference's own code path for the hot path:

    ref.log_mel_spectrogram -> model.encoder(mel) -> loop: model.decoder(tokens, xa, xt_list=[feat]) (no KV cache:
    decoding.py:155-164) -> DecodingTask's own logit filters -> GreedyDecoder.update

i.e. the loop of SURVEY.md Appendix B - `decode()` itself cannot pass the features to a gated x-attn model (F5), so the
loop supplies the one missing argument and uses reference objects for everything else.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref", "whisper"), "/root/reference/whisper")
_ref = None


def locate() -> Optional[str]:
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "__init__.py")) and os.path.isfile(os.path.join(c, "model.py")):
            return c
    return None


def load():
    """Import the reference package as `ref_whisper` (None when it is not available on this box)."""
    global _ref
    if _ref is not None:
        return _ref
    path = locate()
    if path is None:
        return None
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_whisper", os.path.join(path, "__init__.py"),
                                                  submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_whisper"] = mod
    spec.loader.exec_module(mod)
    _ref = mod
    return mod


def build_model(dims: dict, device, gated: bool = True, feat_dim: int = 1024, half: bool = False):
    """Random-init reference model (fp32 master weights - the reference's "half" mode keeps them and casts per call,
    model.py:35-50).  Values do not matter for timing; `torch.empty` parameters are filled."""
    import contextlib
    import io
    ref = load()
    from ref_whisper.model import ModelDimensions, Whisper
    with contextlib.redirect_stdout(io.StringIO()):  # the fork prints "add gated x attn layer" per block
        with torch.device(device):
            model = Whisper(ModelDimensions(**dims), 0.0, False, 256, 1 if gated else 0, feat_dim, 1 if gated else 0)
    model = model.eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("_gate"):
                p.fill_(0.5)
            elif p.dim() >= 2:
                p.normal_(0, 0.02)
        model.decoder.positional_embedding.normal_(0, 0.01)
        if half:  # the upstream notebook's recipe: scalar gates follow the activation dtype
            for name, p in model.named_parameters():
                if name.endswith("_gate"):
                    p.data = p.data.half()
    return ref, model


@torch.no_grad()
def time_hot_path(ref, model, pcm: torch.Tensor, feat: torch.Tensor, sample_len: int, positions, half: bool = False,
                  sync=None):
    """One pass of the reference's hot path over the batch in `pcm` / `feat`, timing the decode loop only at the given
    sampled-token indices (0 .. sample_len-1).  The reference recomputes the whole decoder over all tokens at every
    step, so a step's cost grows with its position; `positions` spread over the whole range keep the extrapolation to
    `sample_len` steps unbiased.  Returns dict(mel_s, enc_s, step_s={pos: seconds}, full_s=extrapolated seconds)."""
    from ref_whisper.decoding import DecodingOptions, DecodingTask
    sync = sync or (lambda: None)
    dev = pcm.device
    opt = DecodingOptions(language="en", task="transcribe", without_timestamps=True, temperature=0.0,
                          sample_len=sample_len, suppress_tokens="-1,50257", suppress_blank=True, fp16=half)
    task = DecodingTask(model, opt)
    task.decoder.reset()
    sync()
    t0 = time.perf_counter()
    mel = ref.log_mel_spectrogram(pcm, n_mels=model.dims.n_mels)
    sync()
    t1 = time.perf_counter()
    if half:
        mel = mel.half()
    xa = model.encoder(mel)
    sync()
    t2 = time.perf_counter()
    n = pcm.shape[0]
    g = torch.Generator().manual_seed(0)
    step_s = {}
    for pos in positions:
        # token history of the right length (values are irrelevant to the cost): prompt + `pos` sampled tokens
        hist = torch.randint(1000, 40000, (n, pos), generator=g)
        tokens = torch.cat([torch.tensor([task.initial_tokens]).repeat(n, 1), hist], dim=1).to(dev)
        sum_lp = torch.zeros(n, device=dev)
        sync()
        s0 = time.perf_counter()
        logits = model.decoder(tokens, xa, xt_list=[feat])          # full recompute, empty kv_cache (decoding.py:155-164)
        logits = logits[:, -1]
        for f in task.logit_filters:
            f.apply(logits, tokens)
        tokens, done = task.decoder.update(tokens, logits, sum_lp)
        bool(done)                                                   # the reference syncs on `completed` every step (:713)
        sync()
        step_s[pos] = time.perf_counter() - s0
    mean_step = sum(step_s.values()) / len(step_s)
    return {"mel_s": t1 - t0, "enc_s": t2 - t1, "step_s": step_s, "mean_step_s": mean_step,
            "full_s": (t1 - t0) + (t2 - t1) + sample_len * mean_step}
