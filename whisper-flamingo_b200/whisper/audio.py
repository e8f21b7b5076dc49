"""Log-mel frontend of the drop-in ``whisper`` package.

Same public surface as reference ``whisper/audio.py`` (constants :13-23, ``pad_or_trim`` :66-89,
``mel_filters`` :92-108, ``log_mel_spectrogram`` :111-161); the computation runs in the fused
CUDA kernel chain of ``csrc/logmel.cu`` behind ``wf_logmel_f32``.  There is no CPU path: CPU
inputs are staged to the current CUDA device and the result is returned on the input's device.
``load_audio`` (:26-63) runs ffmpeg like the reference and reads WAVE files natively where ffmpeg is missing.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _native
from .utils import exact_div

# hard-coded audio hyperparameters (reference audio.py:13-23)
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = exact_div(N_SAMPLES, HOP_LENGTH)  # 3000 frames in a mel spectrogram input
N_VIDEO_FRAMES = 750

N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2  # the initial convolutions has stride 2
FRAMES_PER_SECOND = exact_div(SAMPLE_RATE, HOP_LENGTH)  # 10ms per audio frame
TOKENS_PER_SECOND = exact_div(SAMPLE_RATE, N_SAMPLES_PER_TOKEN)  # 20ms per audio token


def _wav_pcm(file: str):
    """(float32 mono waveform in [-1, 1), sample rate) of a RIFF/WAVE file with integer PCM (8 / 16 / 24 / 32 bit) or
    IEEE float samples; None when the file is something else."""
    import struct
    with open(file, "rb") as fh:
        head = fh.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            return None
        fmt, data = None, None
        while True:
            ck = fh.read(8)
            if len(ck) < 8:
                break
            tag, size = ck[:4], struct.unpack("<I", ck[4:])[0]
            if tag == b"fmt ":
                fmt = fh.read(size)
            elif tag == b"data":
                data = fh.read(size)
            else:
                fh.seek(size, 1)
            if size & 1:
                fh.seek(1, 1)
            if fmt is not None and data is not None:
                break
    if fmt is None or data is None or len(fmt) < 16:
        return None
    code, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if code == 0xFFFE and len(fmt) >= 26:      # WAVE_FORMAT_EXTENSIBLE: the real code opens the sub-format GUID
        code = struct.unpack("<H", fmt[24:26])[0]
    if code == 1 and bits == 8:
        x = (np.frombuffer(data, np.uint8).astype(np.float32) - 128.0) / 128.0
    elif code == 1 and bits == 16:
        x = np.frombuffer(data[: len(data) // 2 * 2], "<i2").astype(np.float32) / 32768.0
    elif code == 1 and bits == 24:
        b = np.frombuffer(data[: len(data) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    elif code == 1 and bits == 32:
        x = np.frombuffer(data[: len(data) // 4 * 4], "<i4").astype(np.float32) / 2147483648.0
    elif code == 3 and bits == 32:
        x = np.frombuffer(data[: len(data) // 4 * 4], "<f4").astype(np.float32)
    elif code == 3 and bits == 64:
        x = np.frombuffer(data[: len(data) // 8 * 8], "<f8").astype(np.float32)
    else:
        return None
    if channels > 1:
        x = x[: len(x) // channels * channels].reshape(-1, channels).mean(axis=1, dtype=np.float32)
    return x, int(rate)


def _resample(x: np.ndarray, rate: int, sr: int) -> np.ndarray:
    """Band-limited rate conversion by a Kaiser-windowed sinc (cut-off at 0.94 of the lower Nyquist, 32 zero
    crossings): one interpolation per output sample, vectorised in blocks.  Not ffmpeg's resampler - only the
    fall-back for hosts without ffmpeg."""
    if rate == sr or len(x) == 0:
        return x.astype(np.float32)
    ratio = sr / rate
    cutoff = 0.94 * min(1.0, ratio)
    half = int(np.ceil(32 / cutoff))
    n_out = int(np.floor(len(x) * ratio))
    xp = np.concatenate([np.zeros(half, np.float32), x.astype(np.float32), np.zeros(half + 1, np.float32)])
    taps = np.arange(-half, half + 1)
    out = np.empty(n_out, np.float32)
    for lo in range(0, n_out, 65536):
        t = np.arange(lo, min(lo + 65536, n_out)) / ratio          # positions in input samples
        base = np.floor(t).astype(np.int64)
        frac = (t - base)[:, None]
        arg = taps[None, :] - frac                                  # distance of every tap from the output instant
        win = np.i0(8.6 * np.sqrt(np.clip(1.0 - (arg / (half + 1)) ** 2, 0.0, None))) / np.i0(8.6)
        h = cutoff * np.sinc(cutoff * arg) * win
        idx = base[:, None] + taps[None, :] + half
        out[lo: lo + len(t)] = (xp[idx] * h).sum(axis=1)
    return out


def load_audio(file: str, sr: int = SAMPLE_RATE):
    """Open an audio file as a mono float32 waveform at ``sr`` Hz (reference audio.py:26-63).

    Like the reference this runs the ``ffmpeg`` CLI (any container / codec, down-mix, resample, 16-bit PCM -> float
    / 32768).  Hosts without ffmpeg still read RIFF/WAVE files natively (PCM or float samples, any rate - resampled by
    a windowed sinc, which is close to but not bit-identical with ffmpeg's)."""
    from subprocess import CalledProcessError, run
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le",
           "-ar", str(sr), "-"]
    try:
        out = run(cmd, capture_output=True, check=True).stdout
        return np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0
    except CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    except FileNotFoundError:
        pass  # no ffmpeg on this host
    wav = _wav_pcm(file)
    if wav is None:
        raise RuntimeError(f"Failed to load audio: ffmpeg is not installed and {file!r} is not a PCM / float WAVE file")
    x, rate = wav
    return _resample(x, rate, sr)


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad (zeros on the right) or trim ``array`` to ``length`` along ``axis``; torch or numpy."""
    n = array.shape[axis]
    if torch.is_tensor(array):
        if n > length:
            array = array.narrow(axis, 0, length)
        elif n < length:
            widths = [0, 0] * array.ndim
            widths[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - n
            array = F.pad(array, widths)
        return array
    if n > length:
        array = np.take(array, range(length), axis=axis)
    elif n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        array = np.pad(array, widths)
    return array


def _slaney_mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """The [n_mels, 201] fp32 filterbank the reference ships as ``assets/mel_filters.npz``
    (``librosa.filters.mel(sr=16000, n_fft=400, n_mels=...)``), regenerated bit-exactly:
    triangle ramps in fp64, rows rounded to fp32, then Slaney area normalisation."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    bins = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = to_hz(np.linspace(to_mel(0.0), to_mel(sr / 2), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - bins[None, :]
    fb = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for m in range(n_mels):
        fb[m] = np.maximum(0, np.minimum(-ramps[m] / width[m], ramps[m + 2] / width[m + 1]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


@lru_cache(maxsize=None)
def _mel_filters_cpu(n_mels: int) -> torch.Tensor:
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(_slaney_mel_filterbank(n_mels))


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int) -> torch.Tensor:
    """Mel filterbank matrix [n_mels, 201] for projecting the STFT power onto mel bins."""
    return _mel_filters_cpu(n_mels).to(device)


def log_mel_spectrogram(
    audio: Union[str, np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    device: Optional[Union[str, torch.device]] = None,
    *,
    per_clip_max: bool = False,
):
    """Log-mel spectrogram of a 16 kHz waveform, shape ``(*, N)`` -> ``(*, n_mels, N // 160)``.

    Semantics follow the reference exactly, including its quirks: anything with a dimension
    of size 80 is returned untouched (audio.py:144), and a batched input is clamped with the
    maximum over the WHOLE batch (audio.py:159).  ``per_clip_max=True`` (keyword-only extension
    used by the batched engine) clamps each clip with its own maximum instead.
    """
    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(np.asarray(audio))
    if 80 in audio.shape:  # already a spectrogram (reference quirk)
        return audio
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    if not torch.cuda.is_available():
        raise _native.WfError("log_mel_spectrogram runs on the B200 CUDA kernels only; no CUDA device is visible")

    src_device = audio.device
    if device is not None:
        audio = audio.to(device)
        src_device = audio.device
    work = audio if audio.is_cuda else audio.to("cuda")
    work = work.to(torch.float32)
    if padding > 0:
        work = F.pad(work, (0, padding))
    lead = work.shape[:-1]
    pcm = work.reshape(-1, work.shape[-1]).contiguous()
    with torch.cuda.device(pcm.device):
        if not _native.logmel_filters_loaded(n_mels):
            _native.logmel_set_filters(n_mels, _mel_filters_cpu(n_mels))
        mode = _native.LOGMEL_PER_CLIP_MAX if per_clip_max else _native.LOGMEL_GLOBAL_MAX
        out = _native.logmel(pcm, n_mels, mode)
    out = out.reshape(*lead, n_mels, out.shape[-1])
    return out if src_device.type == "cuda" else out.to(src_device)
