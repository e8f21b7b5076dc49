"""Oracle: autoregressive decoding (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restatement of the reference's decode loop, ``whisper/decoding.py``:

* ``GreedyDecoder``            decoding.py:276-302  (temperature 0 only)
* ``BeamSearchDecoder``        decoding.py:305-408
* ``SuppressBlank`` / ``SuppressTokens`` / ``ApplyTimestampRules``  decoding.py:427-509
* ``MaximumLikelihoodRanker``  decoding.py:194-217
* ``_main_loop`` + ``run``     decoding.py:688-798

The loop recomputes the whole decoder at every step with an empty KV cache,
exactly as the reference does (decoding.py:155-164), and passes the
feature tensor as ``xt_list=[x_v]`` - the argument the reference's
``PyTorchInference.logits`` drops.  Beam search runs
one audio at a time because the reference only works for ``n_audio == 1``
(decoding.py:743-749).

It is tokenizer-free: everything vocabulary-dependent arrives in ``DecodeSpec``
as plain ids, so the oracle shares no code with the product tokenizer.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import model as om

Tensor = torch.Tensor


@dataclass
class DecodeSpec:
    initial_tokens: Tuple[int, ...]
    eot: int
    sot: int
    no_speech: Optional[int]
    suppress_tokens: Tuple[int, ...]          # decoding.py:619-646 (already expanded)
    blank_tokens: Tuple[int, ...] = ()        # encode(" ") + [eot]; () disables SuppressBlank
    sample_len: int = 224                     # decoding.py:533
    n_ctx: int = 448
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    # timestamp rules (decoding.py:445-509); disabled when without_timestamps
    without_timestamps: bool = True
    timestamp_begin: int = 50364
    no_timestamps: Optional[int] = 50363
    max_initial_timestamp_index: Optional[int] = 50

    @property
    def sample_begin(self) -> int:
        return len(self.initial_tokens)

    @property
    def sot_index(self) -> int:
        return self.initial_tokens.index(self.sot)


@dataclass
class OracleResult:
    tokens: List[int]
    avg_logprob: float
    no_speech_prob: float
    sum_logprob: float


# ----------------------------------------------------------------------------- filters
def apply_filters(spec: DecodeSpec, logits: Tensor, tokens: Tensor) -> None:
    """In-place logit filters in the reference's order (decoding.py:558-574)."""
    if spec.blank_tokens and tokens.shape[1] == spec.sample_begin:      # SuppressBlank :432-434
        logits[:, list(spec.blank_tokens)] = -np.inf
    if spec.suppress_tokens:                                            # SuppressTokens :441-442
        logits[:, list(spec.suppress_tokens)] = -np.inf
    if not spec.without_timestamps:                                     # ApplyTimestampRules :456-509
        tb = spec.timestamp_begin
        if spec.no_timestamps is not None:
            logits[:, spec.no_timestamps] = -np.inf
        for k in range(tokens.shape[0]):
            sampled = tokens[k, spec.sample_begin:]
            seq = sampled.tolist()
            last_ts = len(seq) >= 1 and seq[-1] >= tb
            pen_ts = len(seq) < 2 or seq[-2] >= tb
            if last_ts:
                if pen_ts:
                    logits[k, tb:] = -np.inf
                else:
                    logits[k, : spec.eot] = -np.inf
            ts = sampled[sampled.ge(tb)]
            if ts.numel() > 0:
                if last_ts and not pen_ts:
                    ts_last = ts[-1]
                else:
                    ts_last = ts[-1] + 1
                logits[k, tb:ts_last] = -np.inf
        if tokens.shape[1] == spec.sample_begin:
            logits[:, :tb] = -np.inf
            if spec.max_initial_timestamp_index is not None:
                last_allowed = tb + spec.max_initial_timestamp_index
                logits[:, last_allowed + 1:] = -np.inf
        logprobs = F.log_softmax(logits.float(), dim=-1)
        for k in range(tokens.shape[0]):
            ts_lp = logprobs[k, tb:].logsumexp(dim=-1)
            max_text = logprobs[k, :tb].max()
            if ts_lp > max_text:
                logits[k, :tb] = -np.inf


# ----------------------------------------------------------------------------- token decoders
def greedy_update(spec: DecodeSpec, tokens: Tensor, logits: Tensor, sum_logprobs: Tensor):
    """GreedyDecoder.update at temperature 0, decoding.py:281-297."""
    next_tokens = logits.argmax(dim=-1)
    logprobs = F.log_softmax(logits.float(), dim=-1)
    cur = logprobs[torch.arange(logprobs.shape[0]), next_tokens]
    sum_logprobs += cur * (tokens[:, -1] != spec.eot)
    next_tokens[tokens[:, -1] == spec.eot] = spec.eot
    tokens = torch.cat([tokens, next_tokens[:, None]], dim=-1)
    completed = bool((tokens[:, -1] == spec.eot).all())
    return tokens, completed


class BeamState:
    """BeamSearchDecoder, decoding.py:305-408 (n_audio fixed to 1 by the caller)."""

    def __init__(self, spec: DecodeSpec):
        self.beam = spec.beam_size
        self.eot = spec.eot
        self.patience = spec.patience or 1.0
        self.max_candidates = round(self.beam * self.patience)
        assert self.max_candidates > 0
        self.finished: Optional[List[Dict[tuple, float]]] = None

    def update(self, tokens: Tensor, logits: Tensor, sum_logprobs: Tensor):
        n_audio = tokens.shape[0] // self.beam
        if self.finished is None:
            self.finished = [{} for _ in range(n_audio)]
        logprobs = F.log_softmax(logits.float(), dim=-1)
        next_tokens, source_indices, finished_sequences = [], [], []
        for i in range(n_audio):
            scores, sources, finished = {}, {}, {}
            for j in range(self.beam):
                idx = i * self.beam + j
                prefix = tokens[idx].tolist()
                for logprob, token in zip(*logprobs[idx].topk(self.beam + 1)):
                    new_lp = (sum_logprobs[idx] + logprob).item()
                    seq = tuple(prefix + [token.item()])
                    scores[seq] = new_lp
                    sources[seq] = idx
            saved = 0
            for seq in sorted(scores, key=scores.get, reverse=True):
                if seq[-1] == self.eot:
                    finished[seq] = scores[seq]
                else:
                    sum_logprobs[len(next_tokens)] = scores[seq]
                    next_tokens.append(seq)
                    source_indices.append(sources[seq])
                    saved += 1
                    if saved == self.beam:
                        break
            finished_sequences.append(finished)
        tokens = torch.tensor(next_tokens)
        for prev, new in zip(self.finished, finished_sequences):
            for seq in sorted(new, key=new.get, reverse=True):
                if len(prev) >= self.max_candidates:
                    break
                prev[seq] = new[seq]
        completed = all(len(s) >= self.max_candidates for s in self.finished)
        return tokens, completed

    def finalize(self, preceding: Tensor, sum_logprobs: Tensor):
        sum_logprobs = sum_logprobs.cpu()
        for i, sequences in enumerate(self.finished):
            if len(sequences) < self.beam:
                for j in list(np.argsort(sum_logprobs[i]))[::-1]:
                    seq = preceding[i, j].tolist() + [self.eot]
                    sequences[tuple(seq)] = sum_logprobs[i][j].item()
                    if len(sequences) >= self.beam:
                        break
        toks = [[torch.tensor(s) for s in seqs.keys()] for seqs in self.finished]
        lps = [list(seqs.values()) for seqs in self.finished]
        return toks, lps


def rank(spec: DecodeSpec, tokens: List[List[Tensor]], sum_logprobs: List[List[float]]) -> List[int]:
    """MaximumLikelihoodRanker.rank, decoding.py:203-217."""
    def scores(lps, lens):
        out = []
        for lp, n in zip(lps, lens):
            pen = n if spec.length_penalty is None else ((5 + n) / 6) ** spec.length_penalty
            out.append(lp / pen)
        return out
    lengths = [[len(t) for t in s] for s in tokens]
    return [int(np.argmax(scores(p, l))) for p, l in zip(sum_logprobs, lengths)]


# ----------------------------------------------------------------------------- main loop
@torch.no_grad()
def _run_group(sd, dims: om.Dims, spec: DecodeSpec, xa: Tensor, feat: Optional[Tensor],
               n_group: int, trace: Optional[list] = None):
    """decoding.py:688-765 for the rows in xa (already repeat_interleave'd by n_group)."""
    n_audio = xa.shape[0] // n_group
    tokens = torch.tensor([list(spec.initial_tokens)]).repeat(n_audio, 1).repeat_interleave(n_group, dim=0)
    sum_lp = torch.zeros(tokens.shape[0])
    no_speech = [float("nan")] * tokens.shape[0]
    beam = BeamState(spec) if spec.beam_size is not None else None
    # one feature tensor, or a list of them (multi-"language" gated x-attention, model.py:171-199)
    xt_list = None if feat is None else (list(feat) if isinstance(feat, (list, tuple)) else [feat])
    for i in range(spec.sample_len):
        logits = om.decoder_forward(sd, dims, tokens, xa, xt_list=xt_list)
        if i == 0 and spec.no_speech is not None:
            probs = logits[:, spec.sot_index].float().softmax(dim=-1)
            no_speech = probs[:, spec.no_speech].tolist()
        logits = logits[:, -1]
        if trace is not None:
            trace.append(logits.clone())
        apply_filters(spec, logits, tokens)
        if beam is None:
            tokens, done = greedy_update(spec, tokens, logits, sum_lp)
        else:
            tokens, done = beam.update(tokens, logits, sum_lp)
        if done or tokens.shape[-1] > spec.n_ctx:
            break
    no_speech = no_speech[::n_group]
    tokens = tokens.reshape(n_audio, n_group, -1)
    sum_lp = sum_lp.reshape(n_audio, n_group)
    if beam is None:
        tokens = F.pad(tokens, (0, 1), value=spec.eot)         # GreedyDecoder.finalize :299-302
        cand, lps = tokens, sum_lp.tolist()
    else:
        cand, lps = beam.finalize(tokens, sum_lp)
    cand = [[t[spec.sample_begin: (t == spec.eot).nonzero()[0, 0]] for t in s] for s in cand]
    sel = rank(spec, cand, lps)
    out = []
    for a in range(n_audio):
        toks = cand[a][sel[a]].tolist()
        lp = lps[a][sel[a]]
        out.append(OracleResult(tokens=toks, avg_logprob=lp / (len(toks) + 1),
                                no_speech_prob=no_speech[a], sum_logprob=lp))
    return out


@torch.no_grad()
def decode(sd, dims: om.Dims, spec: DecodeSpec, mel: Tensor, feat: Optional[Tensor] = None,
           act_dtype=torch.float32, trace: Optional[list] = None) -> List[OracleResult]:
    """Encoder + decode loop for a batch of clips.  mel [B,n_mels,3000], feat [B,T_x,bert_dim]."""
    xa = om.encoder_forward(sd, dims, mel, act_dtype=act_dtype)
    if spec.beam_size is None:
        return _run_group(sd, dims, spec, xa, feat, 1, trace)
    out = []
    g = spec.beam_size
    for b in range(xa.shape[0]):  # reference beams: one audio at a time
        if feat is None:
            f = None
        elif isinstance(feat, (list, tuple)):
            f = [x[b:b + 1].repeat_interleave(g, 0) for x in feat]
        else:
            f = feat[b:b + 1].repeat_interleave(g, 0)
        out += _run_group(sd, dims, spec, xa[b:b + 1].repeat_interleave(g, 0), f, g, trace)
    return out


@torch.no_grad()
def detect_language(sd, dims: om.Dims, xa: Tensor, sot: int, language_tokens: Sequence[int],
                    feat=None) -> Tuple[List[int], Tensor]:
    """detect_language, decoding.py:18-77: one decoder pass over [sot], every non-language logit set to -inf,
    argmax + softmax.  Returns (most probable language token per clip, probabilities [B, len(language_tokens)] in the
    order of ``language_tokens``).  ``feat``: feature tensor / list for a gated x-attn model - the reference's
    ``Whisper.logits`` forgets it and crashes there (model.py:374-375), the restatement supplies it."""
    x = torch.tensor([[sot]] * xa.shape[0])
    xt_list = None if feat is None else (list(feat) if isinstance(feat, (list, tuple)) else [feat])
    logits = om.decoder_forward(sd, dims, x, xa, xt_list=xt_list)[:, 0]
    mask = torch.ones(logits.shape[-1], dtype=torch.bool)
    mask[list(language_tokens)] = False
    logits[:, mask] = -np.inf
    best = logits.argmax(dim=-1).tolist()
    probs = logits.softmax(dim=-1)[:, list(language_tokens)]
    return best, probs
