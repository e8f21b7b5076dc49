"""Decoding API of the drop-in ``whisper`` package.

Same surface as reference ``whisper/decoding.py`` (``DecodingOptions`` :80-114, ``DecodingResult``
:117-127, ``DecodingTask`` :512-798, ``decode`` :801-838, ``detect_language`` :18-77) with the
autoregressive loop moved onto the GPU:

* the decoder runs one token per step over engine-owned KV caches (``_engine.DecodeSession``)
  instead of recomputing every position and every cross-attention K/V at each step
  (reference decoding.py:155-164);
* logit filters, argmax / log-softmax, EOT bookkeeping and the "all rows finished" test are device
  kernels inside one replayed CUDA graph - no per-step host sync (reference :704-713);
* ``x_v`` (``[B, T_x, 1024]`` lip/text features, or a list of such tensors) is routed to the gated
  cross-attention as ``xt_list`` - the argument the reference's ``PyTorchInference.logits`` drops
  (SURVEY.md F5); beam search works for any batch size (reference: batch 1 only, F7).

``fp16=True`` (the default) selects the bf16 tensor-core engine, ``fp16=False`` the token-exact
fp32 engine.  ``temperature > 0`` (optionally with ``best_of``) samples on the device with the
Gumbel-max trick and a counter-based RNG seeded from ``torch.initial_seed()``: same distribution as
the reference's ``Categorical(logits / T).sample()``, not the same random stream.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import TYPE_CHECKING, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _engine
from . import _native as nv
from .audio import CHUNK_LENGTH
from .tokenizer import Tokenizer, get_tokenizer
from .utils import compression_ratio

if TYPE_CHECKING:
    from .model import Whisper


@dataclass(frozen=True)
class DecodingOptions:
    task: str = "transcribe"  # "transcribe" (X->X) or "translate" (X->English)
    language: Optional[str] = None  # language of the audio; detected when None
    # sampling
    temperature: float = 0.0
    sample_len: Optional[int] = None  # maximum number of tokens to sample
    best_of: Optional[int] = None  # independent samples when temperature > 0
    beam_size: Optional[int] = None  # beams when temperature == 0
    patience: Optional[float] = None  # beam-search patience (arXiv:2204.05424)
    length_penalty: Optional[float] = None  # GNMT alpha, None = plain length normalisation
    # prompt / prefix (text or token ids)
    prompt: Optional[Union[str, List[int]]] = None
    prefix: Optional[Union[str, List[int]]] = None
    # token suppression: "-1" = the tokenizer's non-speech symbol set
    suppress_tokens: Optional[Union[str, Iterable[int]]] = "-1"
    suppress_blank: bool = True
    # timestamps
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    # half-precision engine (bf16 tensor cores) when True, fp32 engine when False
    fp16: bool = True


@dataclass(frozen=True)
class DecodingResult:
    audio_features: Tensor
    language: str
    language_probs: Optional[Dict[str, float]] = None
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    temperature: float = np.nan
    compression_ratio: float = np.nan


def _as_feature_list(x_v) -> Optional[List[Tensor]]:
    if x_v is None:
        return None
    if torch.is_tensor(x_v):
        return [x_v]
    return list(x_v)


@torch.no_grad()
def detect_language(model: "Whisper", mel: Tensor, tokenizer: Tokenizer = None, x_v=None
                    ) -> Tuple[Tensor, List[dict]]:
    """Most probable language token per clip + the distribution over language tokens
    (reference decoding.py:18-77): one decoder pass over ``[sot]``."""
    if tokenizer is None:
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages)
    if tokenizer.language is None or tokenizer.language_token not in tokenizer.sot_sequence:
        raise ValueError("This model doesn't have language tokens so it can't perform lang id")
    single = mel.ndim == 2
    if single:
        mel = mel.unsqueeze(0)
    if mel.shape[-2:] != (model.dims.n_audio_ctx, model.dims.n_audio_state):
        mel = model.encoder(mel)
    n_audio = mel.shape[0]
    x = torch.tensor([[tokenizer.sot]] * n_audio, device=mel.device)
    logits = model.decoder(x, mel, xt_list=_as_feature_list(x_v))[:, 0]
    mask = torch.ones(logits.shape[-1], dtype=torch.bool, device=logits.device)
    mask[list(tokenizer.all_language_tokens)] = False
    logits[:, mask] = -np.inf
    language_tokens = logits.argmax(dim=-1)
    probs = logits.softmax(dim=-1).cpu()
    language_probs = [
        {c: probs[i, j].item() for j, c in zip(tokenizer.all_language_tokens, tokenizer.all_language_codes)}
        for i in range(n_audio)
    ]
    if single:
        language_tokens, language_probs = language_tokens[0], language_probs[0]
    return language_tokens, language_probs


class MaximumLikelihoodRanker:
    """Pick the hypothesis with the best length-normalised log-probability (reference :194-217)."""

    def __init__(self, length_penalty: Optional[float]):
        self.length_penalty = length_penalty

    def rank(self, tokens: List[List[Sequence[int]]], sum_logprobs: List[List[float]]) -> List[int]:
        def score(lp: float, n: int) -> float:
            pen = n if self.length_penalty is None else ((5 + n) / 6) ** self.length_penalty
            return lp / pen

        return [int(np.argmax([score(lp, len(t)) for lp, t in zip(lps, toks)]))
                for toks, lps in zip(tokens, sum_logprobs)]


class _BeamBook:
    """Host-side bookkeeping of BeamSearchDecoder (reference :305-408), fed with the device top-(k+1)."""

    def __init__(self, beam_size: int, eot: int, patience: Optional[float], n_audio: int):
        self.beam, self.eot = beam_size, eot
        self.max_candidates = round(beam_size * (patience or 1.0))
        assert self.max_candidates > 0, f"Invalid beam size ({beam_size}) or patience ({patience})"
        self.finished: List[Dict[tuple, float]] = [{} for _ in range(n_audio)]
        self._next_id = n_audio  # hypothesis ids; the G prompt rows of audio a start with id a

    def update(self, rows: List[List[int]], row_ids: List[int], sum_lp: np.ndarray, top_lp: np.ndarray,
               top_id: np.ndarray):
        """One step of BeamSearchDecoder.update (reference :337-383).  The reference keys its candidate dictionaries by
        the whole token tuple, which makes identical hypotheses (all beams share the prompt at the first step) collapse
        into one entry; here a hypothesis is identified by an integer id (equal ids <=> equal token sequences), so a
        candidate key is (parent id, token) - same collapsing, same iteration order, no O(length) tuples."""
        next_rows, next_ids, source, new_sum = [], [], [], []
        cand_score = sum_lp.astype(np.float32)[:, None] + top_lp.astype(np.float32)  # fp32 add, as the reference
        k = top_id.shape[1]
        for a in range(len(self.finished)):
            cand: Dict[tuple, tuple] = {}
            for j in range(self.beam):
                r = a * self.beam + j
                rid, sc, ids = row_ids[r], cand_score[r], top_id[r]
                for c in range(k):
                    cand[(rid, int(ids[c]))] = (float(sc[c]), r)  # a later identical hypothesis overwrites (as :349-351)
            kept, done = 0, {}
            for key in sorted(cand, key=lambda q: cand[q][0], reverse=True):
                score, r = cand[key]
                if key[1] == self.eot:
                    done[tuple(rows[r] + [self.eot])] = score
                else:
                    new_sum.append(score)
                    next_rows.append(rows[r] + [key[1]])
                    self._next_id += 1
                    next_ids.append(self._next_id)
                    source.append(r)
                    kept += 1
                    if kept == self.beam:
                        break
            prev = self.finished[a]
            for seq in sorted(done, key=done.get, reverse=True):
                if len(prev) >= self.max_candidates:
                    break
                prev[seq] = done[seq]
        completed = all(len(f) >= self.max_candidates for f in self.finished)
        return next_rows, next_ids, source, np.asarray(new_sum, dtype=np.float32), completed

    def finalize(self, rows: List[List[int]], sum_lp: np.ndarray):
        for a, seqs in enumerate(self.finished):
            if len(seqs) < self.beam:
                order = np.argsort(sum_lp[a * self.beam:(a + 1) * self.beam])[::-1]
                for j in order:
                    seqs[tuple(rows[a * self.beam + j] + [self.eot])] = float(sum_lp[a * self.beam + j])
                    if len(seqs) >= self.beam:
                        break
        return ([[list(s) for s in seqs] for seqs in self.finished],
                [list(seqs.values()) for seqs in self.finished])


class DecodingTask:
    _sample_calls = 0  # advances the sampling RNG stream from call to call (seeded by torch.manual_seed)

    def __init__(self, model: "Whisper", options: DecodingOptions):
        self.model = model
        language = options.language or "en"
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language,
                                  task=options.task)
        self.tokenizer: Tokenizer = tokenizer
        self.options: DecodingOptions = self._verify_options(options)
        self.n_group: int = options.beam_size or options.best_of or 1
        self.n_ctx: int = model.dims.n_text_ctx
        self.sample_len: int = options.sample_len or model.dims.n_text_ctx // 2
        self.sot_sequence: Tuple[int, ...] = tokenizer.sot_sequence
        if self.options.without_timestamps:
            self.sot_sequence = tokenizer.sot_sequence_including_notimestamps
        self.initial_tokens: Tuple[int, ...] = self._get_initial_tokens()
        self.sample_begin: int = len(self.initial_tokens)
        self.sot_index: int = self.initial_tokens.index(tokenizer.sot)
        self.sequence_ranker = MaximumLikelihoodRanker(options.length_penalty)
        # timestamp-rule parameters handed to the sampling kernels: (timestamp_begin, no_timestamps, max_initial)
        self.ts_params = (-1, -1, -1)
        if not options.without_timestamps:
            precision = CHUNK_LENGTH / model.dims.n_audio_ctx  # usually 0.02 s
            max_initial = -1
            if options.max_initial_timestamp:
                max_initial = round(self.options.max_initial_timestamp / precision)
            no_ts = tokenizer.no_timestamps if tokenizer.no_timestamps is not None else -1
            self.ts_params = (tokenizer.timestamp_begin, no_ts, max_initial)

    # ---- option handling (reference :576-646)
    def _verify_options(self, options: DecodingOptions) -> DecodingOptions:
        if options.beam_size is not None and options.best_of is not None:
            raise ValueError("beam_size and best_of can't be given together")
        if options.temperature == 0 and options.best_of is not None:
            raise ValueError("best_of with greedy sampling (T=0) is not compatible")
        if options.patience is not None and options.beam_size is None:
            raise ValueError("patience requires beam_size to be given")
        if options.length_penalty is not None and not (0 <= options.length_penalty <= 1):
            raise ValueError("length_penalty (alpha) should be a value between 0 and 1")
        if options.beam_size is not None and not (1 <= options.beam_size <= 31):
            # the device top-(beam + 1) kernel keeps at most 32 candidates per hypothesis
            raise ValueError(f"beam_size must be between 1 and 31 on this engine, got {options.beam_size}")
        return options

    _UNSET = object()

    def _get_initial_tokens(self, prompt=_UNSET) -> Tuple[int, ...]:
        tokens = list(self.sot_sequence)
        prefix = self.options.prefix
        if prefix:
            ids = self.tokenizer.encode(" " + prefix.strip()) if isinstance(prefix, str) else list(prefix)
            if self.sample_len is not None:
                ids = ids[-(self.n_ctx // 2 - self.sample_len):]
            tokens = tokens + ids
        if prompt is DecodingTask._UNSET:
            prompt = self.options.prompt
        if prompt:
            ids = self.tokenizer.encode(" " + prompt.strip()) if isinstance(prompt, str) else list(prompt)
            tokens = [self.tokenizer.sot_prev] + ids[-(self.n_ctx // 2 - 1):] + tokens
        return tuple(tokens)

    def _get_suppress_tokens(self) -> Tuple[int, ...]:
        ids = self.options.suppress_tokens
        if isinstance(ids, str):
            ids = [int(t) for t in ids.split(",")]
        if ids is None or len(list(ids)) == 0:
            ids = []
        else:
            ids = list(ids)
            if -1 in ids:
                ids = [t for t in ids if t >= 0] + list(self.tokenizer.non_speech_tokens)
        tk = self.tokenizer
        ids += [tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm]
        if tk.no_speech is not None:
            ids.append(tk.no_speech)  # its probability is reported separately
        return tuple(sorted(set(ids)))

    def _masks(self, device) -> Tuple[Tensor, Optional[Tensor]]:
        v = self.model.dims.n_vocab
        suppress = torch.zeros(v, dtype=torch.uint8)
        if self.options.suppress_tokens:
            suppress[list(self._get_suppress_tokens())] = 1
        first = None
        if self.options.suppress_blank:
            first = torch.zeros(v, dtype=torch.uint8)
            first[self.tokenizer.encode(" ") + [self.tokenizer.eot]] = 1
            first = first.to(device)
        return suppress.to(device), first

    # ---- encoder (reference :648-672)
    def _get_audio_features(self, mel: Tensor) -> Tensor:
        dtype = torch.bfloat16 if self.options.fp16 else torch.float32
        dims = self.model.dims
        if mel.shape[-2:] == (dims.n_audio_ctx, dims.n_audio_state):
            return mel.to(dtype)  # already encoded
        return self.model.encoder(mel.to(dtype))

    def _detect_language(self, audio_features: Tensor, tokens: List[List[int]], feats):
        languages = [self.options.language] * audio_features.shape[0]
        lang_probs = None
        if self.options.language is None or self.options.task == "lang_id":
            lang_tokens, lang_probs = self.model.detect_language(audio_features, self.tokenizer, x_v=feats)
            languages = [max(p, key=p.get) for p in lang_probs]
            if self.options.language is None:
                for row, lt in zip(tokens, lang_tokens.tolist()):
                    row[self.sot_index + 1] = lt
        return languages, lang_probs

    # ---- main entry (reference :720-798)
    @torch.no_grad()
    def run(self, mel: Tensor, x_v=None, test_a: bool = False, test_v: bool = False,
            prompts: Optional[Sequence[Sequence[int]]] = None) -> List[DecodingResult]:
        """``prompts`` (extension, used by ``transcribe_batch``): one previous-text token list per clip instead of the
        single ``options.prompt``; after the reference's truncation (:603) they must give rows of equal length."""
        nv.require_cuda(mel)
        tk = self.tokenizer
        n_audio = mel.shape[0]
        feats = _as_feature_list(x_v)
        timing = _engine.PhaseTimer()
        with torch.cuda.device(mel.device):
            timing.mark("start")
            audio_features = self._get_audio_features(mel)
            timing.mark("encoder")
            init_rows = [list(self.initial_tokens) for _ in range(n_audio)]
            if prompts is not None:
                if len(prompts) != n_audio:
                    raise ValueError(f"{len(prompts)} prompts for {n_audio} clips")
                init_rows = [list(self._get_initial_tokens(list(p))) for p in prompts]
                if any(len(r) != self.sample_begin for r in init_rows):
                    raise ValueError("per-clip prompts must give initial token rows of one length "
                                     f"({sorted({len(r) for r in init_rows})} vs {self.sample_begin})")
            languages, language_probs = self._detect_language(audio_features, init_rows, feats)
            if self.options.task == "lang_id":
                return [DecodingResult(audio_features=f, language=l, language_probs=p)
                        for f, l, p in zip(audio_features, languages, language_probs)]
            n_sample = min(self.sample_len, self.n_ctx + 1 - self.sample_begin)
            if n_sample <= 0:
                raise ValueError("the prompt already fills the text context; nothing can be sampled")
            # token capacity of the session in buckets of 64: the prompt of the long-form driver changes length with
            # every window, and a session (K/V arena, step buffers, captured graph) is keyed by its capacity
            t_cap = self.sample_begin + n_sample
            t_cap = min(-(-t_cap // 64) * 64, max(t_cap, self.n_ctx + 1))
            greedy = self.options.beam_size is None
            session = _engine.get_session(self.model.decoder, audio_features, feats, self.n_group, t_cap,
                                          _engine.default_split(n_audio * self.n_group, greedy))
            timing.mark("kv_precompute")
            suppress, suppress_first = self._masks(mel.device)
            no_speech = tk.no_speech if tk.no_speech is not None else -1
            DecodingTask._sample_calls += 1
            session.configure_greedy(self.initial_tokens, self.sot_index, suppress, suppress_first, tk.eot,
                                     no_speech, self.ts_params,
                                     # beam search ignores the temperature (reference :553-559 picks the decoder
                                     # from beam_size alone)
                                     temperature=0.0 if self.options.beam_size else float(self.options.temperature),
                                     seed=(torch.initial_seed() + 0x9E3779B9 * DecodingTask._sample_calls))
            if any(r != list(self.initial_tokens) for r in init_rows):  # detected language tokens differ per clip
                rows = torch.tensor(init_rows, dtype=torch.int32, device=mel.device)
                session.set_initial_rows(rows)
            if self.options.beam_size is None:
                cand, cand_lp, no_speech_probs = self._run_greedy(session, n_sample)
            else:
                cand, cand_lp, no_speech_probs = self._run_beam(session, n_sample, init_rows)
            timing.mark("decode_loop")
            timing.report()
        # slice between the first sampled token and EOT, rank, build results (reference :757-798)
        cand = [[self._trim(seq) for seq in group] for group in cand]
        selected = self.sequence_ranker.rank(cand, cand_lp)
        tokens = [group[i] for i, group in zip(selected, cand)]
        texts = [tk.decode(t).strip() for t in tokens]
        sum_lps = [lp[i] for i, lp in zip(selected, cand_lp)]
        avg_lps = [lp / (len(t) + 1) for t, lp in zip(tokens, sum_lps)]
        fields = (texts, languages, tokens, audio_features, avg_lps, no_speech_probs)
        if len(set(map(len, fields))) != 1:
            raise RuntimeError(f"inconsistent result lengths: {list(map(len, fields))}")
        return [
            DecodingResult(audio_features=f, language=lang, tokens=t, text=text, avg_logprob=alp,
                           no_speech_prob=nsp, temperature=self.options.temperature,
                           compression_ratio=compression_ratio(text))
            for text, lang, t, f, alp, nsp in zip(*fields)
        ]

    def _trim(self, seq: Sequence[int]) -> List[int]:
        seq = list(seq)[self.sample_begin:]
        eot = self.tokenizer.eot
        return seq[: seq.index(eot)] if eot in seq else seq

    def _run_greedy(self, session: "_engine.DecodeSession", n_sample: int):
        session.run_greedy(n_sample)
        toks, lps, nsp = session.results(self.sample_begin + n_sample)
        # GreedyDecoder.finalize pads one EOT so that every row has one (reference :299-302); with best_of the
        # n_group rows of an audio are its independent samples (reference :753-757)
        G, eot = self.n_group, self.tokenizer.eot
        cand = [[toks[a * G + g] + [eot] for g in range(G)] for a in range(len(toks) // G)]
        cand_lp = [[lps[a * G + g] for g in range(G)] for a in range(len(toks) // G)]
        return cand, cand_lp, nsp[::G]

    def _run_beam(self, session: "_engine.DecodeSession", n_sample: int, init_rows: List[List[int]]):
        """BeamSearchDecoder (reference :305-408) with the whole step on the device: decoder pass, top-(G + 1), candidate
        merge per audio, token-history and cache-table permutation are one replayed CUDA graph; the host polls the
        completion flag every 8 steps and assembles the candidate lists once at the end (finalize, :388-408)."""
        import os
        if os.environ.get("WF_BEAM_HOST", "0") == "1":
            return self._run_beam_host(session, n_sample, init_rows)
        G, tk = self.n_group, self.tokenizer
        max_candidates = round(G * (self.options.patience or 1.0))
        assert max_candidates > 0, f"Invalid beam size ({G}) or patience ({self.options.patience})"
        if max_candidates > 32:
            raise ValueError(f"beam_size * patience = {max_candidates} finished candidates per audio; this engine keeps "
                             f"at most 32")
        session.configure_beam(max_candidates)
        session.run_beam(n_sample, self.n_ctx)
        # one device -> host read of the whole search state
        L = int(session.state[0].item()) + 1          # tokens per live hypothesis
        rows = session.tokens[:, :L].cpu().tolist()
        sum_lp = session.sum_logprobs.cpu().numpy()
        n_fin = session.n_fin.cpu().tolist()
        fin_len = session.fin_len.cpu().tolist()
        fin_score = session.fin_score.cpu().tolist()
        fin_tok = session.fin_tokens.cpu()
        cand, cand_lp = [], []
        for a in range(session.B):
            seqs = {tuple(fin_tok[a, m, : fin_len[a][m]].tolist()): fin_score[a][m] for m in range(n_fin[a])}
            if len(seqs) < G:  # top up with the best live hypotheses (reference :393-400)
                order = np.argsort(sum_lp[a * G:(a + 1) * G])[::-1]
                for j in order:
                    seqs[tuple(rows[a * G + j] + [tk.eot])] = float(sum_lp[a * G + j])
                    if len(seqs) >= G:
                        break
            cand.append([list(q) for q in seqs])
            cand_lp.append(list(seqs.values()))
        nsp = session.no_speech_prob.cpu().tolist()[::G]
        return cand, cand_lp, nsp

    def _run_beam_host(self, session: "_engine.DecodeSession", n_sample: int, init_rows: List[List[int]]):
        """Round-1 form: the decoder pass and the top-k on the device, the candidate merge in Python (two device -> host
        syncs and three uploads per step).  WF_BEAM_HOST=1; kept as the A/B comparator of the device-side search."""
        G, tk, n_init = self.n_group, self.tokenizer, self.sample_begin
        R = session.R
        book = _BeamBook(G, tk.eot, self.options.patience, session.B)
        rows = [list(init_rows[r // G]) for r in range(R)]
        row_ids = [r // G for r in range(R)]
        sum_lp = np.zeros(R, dtype=np.float32)
        suppress, suppress_first, eot, no_speech, ts = session._sampler[:5]
        for _ in range(n_init - 1):  # feed the prompt; the sampler only records no_speech_prob here
            session._greedy_step()
        k = G + 1
        vals = torch.empty((R, k), dtype=torch.float32, device=session.dev)
        idx = torch.empty((R, k), dtype=torch.int32, device=session.dev)
        identity = list(range(R))
        pos = n_init - 1
        for i in range(n_sample):
            session.forward_at(pos)  # one replay of the captured decoder pass
            if pos == self.sot_index and no_speech >= 0:
                # SOT is the last prompt token (English-only vocabularies): take no_speech_prob from this pass.
                # The sampler's other outputs (tokens[:, pos + 1], device sum_logprobs) are overwritten / unused here.
                nv.sample_greedy(session.logits, session.p.n_vocab, suppress, suppress_first, session.tokens,
                                 session.state, session.sum_logprobs, session.no_speech_prob, eot, no_speech, ts)
            nv.topk_logprobs(session.logits, session.p.n_vocab, suppress, suppress_first, session.tokens,
                             n_init, len(rows[0]), eot, ts, k, vals, idx)
            rows, row_ids, source, sum_lp, completed = book.update(rows, row_ids, sum_lp, vals.cpu().numpy(),
                                                                  idx.cpu().numpy())
            pos += 1
            # token history on the device: row r continues row source[r] and appends its new token
            src_dev = torch.tensor(source, dtype=torch.int64, device=session.dev)
            L = len(rows[0])
            if source != identity:
                session.tokens[:, : L - 1] = session.tokens[:, : L - 1].index_select(0, src_dev)
                session.reorder_self_kv(src_dev, pos)
            session.tokens[:, L - 1] = torch.tensor([row[-1] for row in rows], dtype=torch.int32, device=session.dev)
            if completed or len(rows[0]) > self.n_ctx:
                break
        cand, cand_lp = book.finalize(rows, sum_lp)
        nsp = session.no_speech_prob.cpu().tolist()[::G]
        return cand, cand_lp, nsp


@torch.no_grad()
def decode(model: "Whisper", mel: Tensor, options: DecodingOptions = DecodingOptions(), x_v=None,
           test_v: bool = False, test_a: bool = False, prompts: Optional[Sequence[Sequence[int]]] = None,
           **kwargs) -> Union[DecodingResult, List[DecodingResult]]:
    """Decode 30-second segment(s) given as mel spectrogram(s) ``(n_mels, 3000)`` or ``(*, n_mels, 3000)``.

    ``x_v``: optional feature tensor ``(*, T_x, bert_dim)`` (or list of tensors) for a gated x-attn model.
    ``test_a`` / ``test_v`` are accepted for signature compatibility (dead arguments in the reference).
    """
    single = mel.ndim == 2
    if single:
        mel = mel.unsqueeze(0)
        if torch.is_tensor(x_v) and x_v.ndim == 2:
            x_v = x_v.unsqueeze(0)
    if kwargs:
        options = replace(options, **kwargs)
    result = DecodingTask(model, options).run(mel, x_v, test_a, test_v, prompts=prompts)
    return result[0] if single else result
