"""Long-form transcription: 30-second sliding window over the log-mel of a whole recording.

Same contract as reference ``whisper/transcribe.py:38-383`` (``transcribe(model, audio, *, verbose, temperature,
compression_ratio_threshold, logprob_threshold, no_speech_threshold, condition_on_previous_text,
initial_prompt, word_timestamps, ..., **decode_options) -> {"text", "segments", "language"}``): every window
is decoded by ``whisper.decode`` on the CUDA engine (timestamp grammar on the device), with the
temperature-fallback ladder, no-speech skipping, timestamp-token segmentation and prompt conditioning
done on the host exactly as the reference does.  ``x_v`` (features for a gated x-attn model, aligned to the
whole recording at 25 frames / s) is an extension: the slice matching each window is passed to ``decode``.

Word-level timestamps (``word_timestamps=True``) are out of scope (the DTW / median-filter path of
``whisper/timing.py`` is removed by the north star) and raise ``NotImplementedError``.
"""
from __future__ import annotations

from typing import TYPE_CHECKING, List, Optional, Tuple, Union

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
    if word_timestamps:
        raise NotImplementedError("word-level timestamps (timing.py / triton_ops.py) are outside the B200 hot path")
    fp16 = decode_options.get("fp16", True)
    dtype = torch.bfloat16 if fp16 else torch.float32
    device = model.device
    if device.type != "cuda":
        raise RuntimeError("transcribe needs the model on a CUDA device (no CPU fallback exists)")

    # whole-recording log-mel with 30 s of trailing silence so that every window can be sliced at full width
    mel = log_mel_spectrogram(audio, model.dims.n_mels, padding=N_SAMPLES, device=device)
    content_frames = mel.shape[-1] - N_FRAMES

    if decode_options.get("language") is None:
        if not model.is_multilingual:
            decode_options["language"] = "en"
        else:
            if verbose:
                print("Detecting language using up to the first 30 seconds. Use `--language` to specify the language")
            first = pad_or_trim(mel, N_FRAMES).to(dtype)
            feats0 = None if x_v is None else x_v[: N_FRAMES // 4][None].to(device)
            _, probs = model.detect_language(first, x_v=feats0)
            decode_options["language"] = max(probs, key=probs.get)
            if verbose is not None:
                print(f"Detected language: {LANGUAGES[decode_options['language']].title()}")
    language: str = decode_options["language"]
    task: str = decode_options.get("task", "transcribe")
    tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language, task=task)
    temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)

    def decode_with_fallback(segment: torch.Tensor, feats) -> DecodingResult:
        result = None
        for t in temperatures:
            kwargs = dict(decode_options)
            if t > 0:  # beam search is a temperature-0 procedure
                kwargs.pop("beam_size", None)
                kwargs.pop("patience", None)
            else:
                kwargs.pop("best_of", None)
            result = model.decode(segment, DecodingOptions(**kwargs, temperature=t), x_v=feats)
            if not _needs_fallback(result, compression_ratio_threshold, logprob_threshold, no_speech_threshold):
                break
        return result

    input_stride = exact_div(N_FRAMES, model.dims.n_audio_ctx)  # mel frames per encoder position: 2
    time_precision = input_stride * HOP_LENGTH / SAMPLE_RATE      # seconds per timestamp token step: 0.02
    all_tokens: List[int] = []
    all_segments: List[dict] = []
    prompt_reset_since = 0
    initial_prompt_tokens: List[int] = []
    if initial_prompt is not None:
        initial_prompt_tokens = tokenizer.encode(" " + initial_prompt.strip())
        all_tokens.extend(initial_prompt_tokens)

    seek = 0
    while seek < content_frames:
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        segment_size = min(N_FRAMES, content_frames - seek)
        segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
        window = pad_or_trim(mel[:, seek: seek + N_FRAMES], N_FRAMES).to(dtype)
        feats = None
        if x_v is not None:  # features run at 25 fps = one per 4 mel frames
            f0 = seek // 4
            feats = pad_or_trim(x_v[f0: f0 + N_FRAMES // 4].to(device), N_FRAMES // 4, axis=0)

        decode_options["prompt"] = all_tokens[prompt_reset_since:]
        result = decode_with_fallback(window, feats)
        tokens = list(result.tokens)

        if no_speech_threshold is not None:
            skip = result.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                skip = False  # confident text despite a high no-speech probability
            if skip:
                seek += segment_size
                continue

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
                current.append(make(time_offset + (piece[0] - tb) * time_precision,
                                    time_offset + (piece[-1] - tb) * time_precision, piece))
            if not single_ending:  # drop the unfinished tail: resume at the last closed timestamp
                advance = (tokens[spans[-1][1] - 1] - tb) * input_stride
        else:
            duration = segment_duration
            stamps = [t for t in tokens if t >= tb]
            if stamps and stamps[-1] != tb:
                duration = (stamps[-1] - tb) * time_precision
            current.append(make(time_offset, time_offset + duration, tokens))
        seek += advance

        if verbose:
            for seg in current:
                print(f"[{seg['start']:8.2f} --> {seg['end']:8.2f}] {seg['text']}")
        for seg in current:  # instantaneous or empty segments carry no text
            if seg["start"] == seg["end"] or seg["text"].strip() == "":
                seg["text"], seg["tokens"], seg["words"] = "", [], []
        for seg in current:
            all_segments.append({"id": len(all_segments), **seg})
        all_tokens.extend(t for seg in current for t in seg["tokens"])
        if not condition_on_previous_text or result.temperature > 0.5:
            prompt_reset_since = len(all_tokens)  # do not condition on text sampled at a high temperature

    return dict(text=tokenizer.decode(all_tokens[len(initial_prompt_tokens):]), segments=all_segments,
                language=language)
