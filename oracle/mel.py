"""Oracle: log-mel frontend (TEST INFRASTRUCTURE, see oracle/__init__.py).

numpy restatement of reference ``whisper/audio.py``:

* constants                   audio.py:13-23
* ``pad_or_trim``             audio.py:66-89
* ``mel_filters``             audio.py:92-108   (the .npz there was produced by
                              ``librosa.filters.mel(sr=16000, n_fft=400, n_mels=N)``;
                              ``mel_filterbank`` below regenerates it bit-exactly -
                              sha256 prefixes pinned in tests/test_oracle_cpu.py)
* ``log_mel_spectrogram``     audio.py:111-161

The STFT (``torch.stft(audio, 400, 160, window=hann_window(400),
return_complex=True)``, audio.py:151-153) is restated literally: centre
reflect-pad 200, 3001 frames of 400 at hop 160, periodic Hann, one-sided
rDFT (201 bins), last frame dropped, ``|X|^2``.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000


def pad_or_trim(array: np.ndarray, length: int = N_SAMPLES, axis: int = -1) -> np.ndarray:
    """audio.py:66-89 (numpy branch)."""
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """Slaney-style triangular mel filterbank == the array stored in the
    reference's ``assets/mel_filters.npz`` (audio.py:92-108), fp32 [n_mels, 201].

    Arithmetic order matters for bit-exactness: ramps in fp64, rows rounded to
    fp32, *then* scaled by the fp64 Slaney normaliser and rounded again.
    """
    assert n_mels in (80, 128), f"Unsupported n_mels: {n_mels}"
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz,
                        min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep,
                        f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w


def stft_power(audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """|STFT|^2 with the last frame dropped: (..., N) -> (..., 201, n_frames).

    audio.py:151-153.  ``dtype`` is the arithmetic type (fp64 = ground truth,
    fp32 = same precision class as the reference's pocketfft/cuFFT path).
    """
    x = np.asarray(audio, dtype=dtype)
    pad = N_FFT // 2
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    n_frames = 1 + (xp.shape[-1] - N_FFT) // HOP_LENGTH
    n = np.arange(N_FFT, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)).astype(dtype)  # periodic Hann
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange(n_frames)[:, None]
    frames = xp[..., idx] * window  # (..., n_frames, 400)
    spec = np.fft.rfft(frames, axis=-1)  # (..., n_frames, 201)
    power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2)
    power = np.swapaxes(power, -1, -2)  # (..., 201, n_frames)
    return power[..., :-1]


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, padding: int = 0,
                        dtype=np.float64, per_clip_max: bool = False) -> np.ndarray:
    """audio.py:111-161.  Returns fp32 (..., n_mels, n_frames).

    ``per_clip_max=False`` is the reference semantics: the ``max - 8`` clamp uses
    the max over the WHOLE tensor (audio.py:159).  ``True`` is the
    per-clip variant the batched engine uses (every reference caller invokes the
    function one clip at a time, where the two coincide).
    """
    audio = np.asarray(audio)
    if 80 in audio.shape:  # audio.py:144 passthrough quirk
        return audio
    if padding > 0:
        audio = np.pad(audio, [(0, 0)] * (audio.ndim - 1) + [(0, padding)])
    power = stft_power(audio, dtype=dtype)
    filters = mel_filterbank(n_mels).astype(dtype)
    mel_spec = filters @ power
    log_spec = np.log10(np.maximum(mel_spec, 1e-10))
    if per_clip_max and log_spec.ndim > 2:
        mx = log_spec.max(axis=(-1, -2), keepdims=True)
    else:
        mx = log_spec.max()
    log_spec = np.maximum(log_spec, mx - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec.astype(np.float32)


def chirp_kat(n: int = N_SAMPLES) -> np.ndarray:
    """RNG-free known-answer input of SURVEY.md section 8(c):
    x(t) = 0.5 sin(2pi(100 t + 120 t^2)) + 0.05 sin(2pi 3000 t), fp64 -> fp32."""
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    x = 0.5 * np.sin(2 * np.pi * (100 * t + 120 * t * t)) + 0.05 * np.sin(2 * np.pi * 3000 * t)
    return x.astype(np.float32)
