"""Long-form transcription: 30-second sliding window over the log-mel of a whole recording.

Same contract as reference ``whisper/transcribe.py:38-383`` (``transcribe(model, audio, *, verbose, temperature,
compression_ratio_threshold, logprob_threshold, no_speech_threshold, condition_on_previous_text,
initial_prompt, word_timestamps, ..., **decode_options) -> {"text", "segments", "language"}``): every window
is decoded by ``whisper.decode`` on the CUDA engine (timestamp grammar on the device), with the
temperature-fallback ladder, no-speech skipping, timestamp-token segmentation and prompt conditioning
done on the host exactly as the reference does.  ``x_v`` (features for a gated x-attn model, aligned to the
whole recording at 25 frames / s) is an extension: the slice matching each window is passed to ``decode``.

``transcribe_batch`` is the B200 form of the same driver: the current window of MANY recordings is decoded as one
batch per round, each recording keeping the reference's state machine.

Word-level timestamps (``word_timestamps=True``) are out of scope (the DTW / median-filter path of
``whisper/timing.py`` is removed by the north star) and raise ``NotImplementedError``.
"""
from __future__ import annotations

from typing import TYPE_CHECKING, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult
from .tokenizer import LANGUAGES, get_tokenizer
from .utils import exact_div

if TYPE_CHECKING:
    from .model import Whisper

VIDEO_FRAMES_PER_SECOND = 25


def _needs_fallback(result: DecodingResult, compression_ratio_threshold, logprob_threshold, no_speech_threshold) -> bool:
    retry = False
    if compression_ratio_threshold is not None and result.compression_ratio > compression_ratio_threshold:
        retry = True  # too repetitive
    if logprob_threshold is not None and result.avg_logprob < logprob_threshold:
        retry = True  # too unlikely
    if no_speech_threshold is not None and result.no_speech_prob > no_speech_threshold:
        retry = False  # silence: nothing better to be had at a higher temperature
    return retry


def _cut_segments(tokens: List[int], timestamp_begin: int) -> Tuple[List[Tuple[int, int]], bool]:
    """Token ranges [lo, hi) delimited by consecutive timestamp pairs; second value: the window ended on a single
    timestamp (nothing spoken after it)."""
    is_ts = [t >= timestamp_begin for t in tokens]
    single_ending = len(tokens) >= 2 and (not is_ts[-2]) and is_ts[-1]
    if len(tokens) == 1:
        single_ending = False  # reference: tokens[-2:] == [False, True] needs two entries
    cuts = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
    if not cuts:
        return [], single_ending
    if single_ending:
        cuts.append(len(tokens))
    spans, lo = [], 0
    for hi in cuts:
        spans.append((lo, hi))
        lo = hi
    return spans, single_ending


class _Recording:
    """Window-by-window state of one recording (reference transcribe.py:234-377): seek position, accumulated tokens
    and segments, prompt window.  ``window()`` hands out the next 30-s slice, ``consume()`` books a decoding result."""

    def __init__(self, model, mel, x_v, tokenizer, initial_prompt, condition_on_previous_text, no_speech_threshold,
                 logprob_threshold, dtype, verbose):
        self.model, self.mel, self.x_v, self.tokenizer = model, mel, x_v, tokenizer
        self.condition, self.no_speech_threshold, self.logprob_threshold = (condition_on_previous_text,
                                                                            no_speech_threshold, logprob_threshold)
        self.dtype, self.verbose = dtype, verbose
        self.content_frames = mel.shape[-1] - N_FRAMES
        self.input_stride = exact_div(N_FRAMES, model.dims.n_audio_ctx)     # mel frames per encoder position: 2
        self.time_precision = self.input_stride * HOP_LENGTH / SAMPLE_RATE  # seconds per timestamp token step: 0.02
        self.all_tokens: List[int] = []
        self.all_segments: List[dict] = []
        self.prompt_reset_since = 0
        self.initial_prompt_tokens: List[int] = []
        if initial_prompt is not None:
            self.initial_prompt_tokens = tokenizer.encode(" " + initial_prompt.strip())
            self.all_tokens.extend(self.initial_prompt_tokens)
        self.seek = 0

    @property
    def done(self) -> bool:
        return self.seek >= self.content_frames

    def prompt(self) -> List[int]:
        return self.all_tokens[self.prompt_reset_since:]

    def window(self):
        """(mel window [n_mels, 3000], feature slice or None) at the current seek position."""
        device = self.mel.device
        win = pad_or_trim(self.mel[:, self.seek: self.seek + N_FRAMES], N_FRAMES).to(self.dtype)
        feats = None
        if self.x_v is not None:  # features run at 25 fps = one per 4 mel frames
            f0 = self.seek // 4
            feats = pad_or_trim(self.x_v[f0: f0 + N_FRAMES // 4].to(device), N_FRAMES // 4, axis=0)
        return win, feats

    def consume(self, result: DecodingResult) -> None:
        tokenizer, seek = self.tokenizer, self.seek
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        segment_size = min(N_FRAMES, self.content_frames - seek)
        segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
        tokens = list(result.tokens)

        if self.no_speech_threshold is not None:
            skip = result.no_speech_prob > self.no_speech_threshold
            if self.logprob_threshold is not None and result.avg_logprob > self.logprob_threshold:
                skip = False  # confident text despite a high no-speech probability
            if skip:
                self.seek += segment_size
                return

        def make(start: float, end: float, toks: List[int]) -> dict:
            return {"seek": seek, "start": start, "end": end,
                    "text": tokenizer.decode([t for t in toks if t < tokenizer.eot]), "tokens": toks,
                    "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                    "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob}

        tb = tokenizer.timestamp_begin
        current: List[dict] = []
        spans, single_ending = _cut_segments(tokens, tb)
        advance = segment_size
        if spans:
            for lo, hi in spans:
                piece = tokens[lo:hi]
                current.append(make(time_offset + (piece[0] - tb) * self.time_precision,
                                    time_offset + (piece[-1] - tb) * self.time_precision, piece))
            if not single_ending:  # drop the unfinished tail: resume at the last closed timestamp
                advance = (tokens[spans[-1][1] - 1] - tb) * self.input_stride
        else:
            duration = segment_duration
            stamps = [t for t in tokens if t >= tb]
            if stamps and stamps[-1] != tb:
                duration = (stamps[-1] - tb) * self.time_precision
            current.append(make(time_offset, time_offset + duration, tokens))
        self.seek += advance

        if self.verbose:
            for seg in current:
                print(f"[{seg['start']:8.2f} --> {seg['end']:8.2f}] {seg['text']}")
        for seg in current:  # instantaneous or empty segments carry no text
            if seg["start"] == seg["end"] or seg["text"].strip() == "":
                seg["text"], seg["tokens"], seg["words"] = "", [], []
        for seg in current:
            self.all_segments.append({"id": len(self.all_segments), **seg})
        self.all_tokens.extend(t for seg in current for t in seg["tokens"])
        if not self.condition or result.temperature > 0.5:
            self.prompt_reset_since = len(self.all_tokens)  # do not condition on text sampled at a high temperature

    def result(self, language: str) -> dict:
        return dict(text=self.tokenizer.decode(self.all_tokens[len(self.initial_prompt_tokens):]),
                    segments=self.all_segments, language=language)


def _temperature_kwargs(decode_options: dict, t: float) -> dict:
    kwargs = dict(decode_options)
    if t > 0:  # beam search is a temperature-0 procedure
        kwargs.pop("beam_size", None)
        kwargs.pop("patience", None)
    else:
        kwargs.pop("best_of", None)
    return kwargs


def transcribe(
    model: "Whisper",
    audio: Union[str, np.ndarray, torch.Tensor],
    *,
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    word_timestamps: bool = False,
    prepend_punctuations: str = "\"'“¿([{-",
    append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
    video: bool = False,
    x_v: Optional[torch.Tensor] = None,
    **decode_options,
):
    """One recording, window after window (the reference's loop); see ``transcribe_batch`` for many recordings."""
    return transcribe_batch(model, [audio], verbose=verbose, temperature=temperature,
                            compression_ratio_threshold=compression_ratio_threshold,
                            logprob_threshold=logprob_threshold, no_speech_threshold=no_speech_threshold,
                            condition_on_previous_text=condition_on_previous_text, initial_prompt=initial_prompt,
                            word_timestamps=word_timestamps, x_v=None if x_v is None else [x_v], **decode_options)[0]


def _bucket_rows(n: int, batch_bucket: int, max_batch: int) -> int:
    """Rows a decode call of n windows is padded to: the next multiple of `batch_bucket`, a power of two below one
    bucket (a single recording stays a single row), never more than `max_batch`; `batch_bucket` <= 1 turns it off."""
    if batch_bucket <= 1 or n <= 0:
        return n
    bucket = -(-n // batch_bucket) * batch_bucket if n >= batch_bucket else 1 << (n - 1).bit_length()
    return bucket if bucket <= max_batch else n


def transcribe_batch(
    model: "Whisper",
    audios: Sequence[Union[str, np.ndarray, torch.Tensor]],
    *,
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    word_timestamps: bool = False,
    x_v: Optional[Sequence[Optional[torch.Tensor]]] = None,
    max_batch: int = 128,
    batch_bucket: int = 16,
    **decode_options,
) -> List[dict]:
    """Long-form transcription of MANY recordings at once: in every round the current 30-s window of each unfinished
    recording is decoded as ONE batch on the engine (the reference decodes one window of one file at a time,
    transcribe.py:234-377).  Each recording keeps the reference's own state machine - seek, timestamp segmentation,
    no-speech skipping, prompt conditioning, temperature fallback - so the result of every recording equals what
    ``transcribe`` returns for it alone.

    Windows are grouped by (language, length of the previous-text prompt): a decode session shares one prompt length
    (per-clip prompt *contents* are free).  Without ``condition_on_previous_text`` every window of a language is one
    group.  Rows that fail the compression / log-probability thresholds are decoded again, as a sub-batch, at the next
    temperature of the ladder.  ``x_v``: one feature tensor (25 frames / s over the whole recording) per recording."""
    if word_timestamps:
        raise NotImplementedError("word-level timestamps (timing.py / triton_ops.py) are outside the B200 hot path")
    fp16 = decode_options.get("fp16", True)
    dtype = torch.bfloat16 if fp16 else torch.float32
    device = model.device
    if device.type != "cuda":
        raise RuntimeError("transcribe needs the model on a CUDA device (no CPU fallback exists)")
    audios = list(audios)
    feats_all = list(x_v) if x_v is not None else [None] * len(audios)
    if len(feats_all) != len(audios):
        raise ValueError(f"{len(feats_all)} feature tensors for {len(audios)} recordings")

    # whole-recording log-mel with 30 s of trailing silence so that every window can be sliced at full width
    mels = [log_mel_spectrogram(a, model.dims.n_mels, padding=N_SAMPLES, device=device) for a in audios]

    # ---- language per recording (reference :126-140: detected on the first 30 s unless given)
    languages: List[str] = [decode_options.get("language")] * len(audios)
    if decode_options.get("language") is None:
        if not model.is_multilingual:
            languages = ["en"] * len(audios)
        else:
            if verbose:
                print("Detecting language using up to the first 30 seconds. Use `--language` to specify the language")
            for lo in range(0, len(audios), max_batch):
                idx = range(lo, min(lo + max_batch, len(audios)))
                first = torch.stack([pad_or_trim(mels[i], N_FRAMES).to(dtype) for i in idx])
                f0 = None
                if all(feats_all[i] is not None for i in idx):
                    f0 = torch.stack([pad_or_trim(feats_all[i][: N_FRAMES // 4].to(device), N_FRAMES // 4, axis=0)
                                      for i in idx])
                _, probs = model.detect_language(first, x_v=f0)
                for i, p in zip(idx, probs):
                    languages[i] = max(p, key=p.get)
                    if verbose is not None:
                        print(f"Detected language: {LANGUAGES[languages[i]].title()}")
    task: str = decode_options.get("task", "transcribe")
    temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)
    n_prompt_max = model.dims.n_text_ctx // 2 - 1     # the reference keeps the last 223 prompt tokens (decoding.py:603)

    recs = [_Recording(model, mels[i], feats_all[i],
                       get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=languages[i],
                                     task=task),
                       initial_prompt, condition_on_previous_text, no_speech_threshold, logprob_threshold, dtype, verbose)
            for i in range(len(audios))]

    def decode_group(idx: List[int], language: str) -> None:
        """One window of every recording in idx (same language, same prompt length) through the fallback ladder."""
        wins, fts = zip(*(recs[i].window() for i in idx))
        mel_b = torch.stack(list(wins))
        feat_b = torch.stack(list(fts)) if fts[0] is not None else None
        prompts = [recs[i].prompt() for i in idx]
        pending = list(range(len(idx)))
        results: List[Optional[DecodingResult]] = [None] * len(idx)
        for t in temperatures:
            kwargs = _temperature_kwargs(decode_options, t)
            kwargs["language"] = language
            kwargs["prompt"] = prompts[pending[0]]
            # Batch sizes are bucketed (multiples of `batch_bucket`, powers of two below it; padded with repeats of the last row whose results
            # are dropped): the batch shrinks as recordings end and every fallback rung has its own row count - each
            # distinct size would otherwise build a new decode session (arena + CUDA-graph capture).  Rows are
            # independent of each other, so the real rows decode exactly as they would alone.
            rows = list(pending)
            rows = rows + [rows[-1]] * (_bucket_rows(len(rows), batch_bucket, max_batch) - len(rows))
            whole = rows == list(range(len(idx)))
            sel = torch.tensor(rows, device=device)
            out = model.decode(mel_b if whole else mel_b.index_select(0, sel), DecodingOptions(**kwargs, temperature=t),
                               x_v=None if feat_b is None else (feat_b if whole else feat_b.index_select(0, sel)),
                               prompts=[prompts[j] for j in rows])
            out = out[: len(pending)]
            again = []
            for j, r in zip(pending, out):
                results[j] = r
                if _needs_fallback(r, compression_ratio_threshold, logprob_threshold, no_speech_threshold):
                    again.append(j)
            pending = again
            if not pending:
                break
        for i, r in zip(idx, results):
            recs[i].consume(r)

    while True:
        active = [i for i, r in enumerate(recs) if not r.done]
        if not active:
            break
        groups: dict = {}
        for i in active:
            groups.setdefault((languages[i], min(len(recs[i].prompt()), n_prompt_max)), []).append(i)
        for (language, _), idx in groups.items():
            for lo in range(0, len(idx), max_batch):
                decode_group(idx[lo: lo + max_batch], language)

    return [r.result(languages[i]) for i, r in enumerate(recs)]
