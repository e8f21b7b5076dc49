"""BASELINE config 5: log-mel frontend sweep, 1 .. 4096 clips of 30 s, 80 and 128 mel bins.

For every batch: the fused libwf kernel chain (through whisper.log_mel_spectrogram) against the reference's algorithm
(whisper/audio.py:147-160: torch.stft -> |X|^2 -> filterbank matmul -> log10 -> max-8 clamp -> (x+4)/4) executed by
PyTorch on the SAME GPU, both reported as clips/s and as GB/s of algorithmic bytes (PCM in + fp32 log-mel out), plus
the max abs difference between the two.  The torch path below is a benchmark comparator, not product code.
usage: python tools/mel_sweep.py > profiles/r01_mel_sweep.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-flamingo_b200"))
import torch
import whisper
from whisper.audio import mel_filters


def torch_stft_logmel(pcm, filters):
    window = torch.hann_window(400, device=pcm.device)
    stft = torch.stft(pcm, 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    spec = torch.clamp(filters @ mag, min=1e-10).log10()
    spec = torch.maximum(spec, spec.amax(dim=(-2, -1), keepdim=True) - 8.0)     # per clip, as the engine uses it
    return (spec + 4.0) / 4.0


def timeit(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    print("log-mel sweep on", torch.cuda.get_device_name(0))
    print(f"{'clips':>6s} {'mels':>5s} | {'libwf us':>10s} {'clips/s':>10s} {'GB/s':>7s} | {'torch.stft us':>13s} {'clips/s':>10s} "
          f"{'GB/s':>7s} | {'speed-up':>8s} {'max |diff|':>10s}")
    for n_mels in (80, 128):
        filt = mel_filters("cuda", n_mels)
        for B in (1, 2, 4, 16, 64, 256, 1024, 4096):
            pcm = torch.randn(B, 480000, device="cuda") * 0.1
            by = B * (480000 * 4 + n_mels * 3000 * 4)
            iters = 20 if B <= 64 else 5
            t_wf = timeit(lambda: whisper.log_mel_spectrogram(pcm, n_mels=n_mels, per_clip_max=True), iters)
            step = min(B, 256)          # the torch path materialises complex64 [B, 201, 3001] + its square: bounded slices
            def ref():
                return [torch_stft_logmel(pcm[i:i + step], filt) for i in range(0, B, step)]
            t_ref = timeit(ref, max(2, iters // 4))
            a = whisper.log_mel_spectrogram(pcm[:step], n_mels=n_mels, per_clip_max=True)
            diff = float((a - torch_stft_logmel(pcm[:step], filt)).abs().max())
            print(f"{B:6d} {n_mels:5d} | {t_wf:10.1f} {B / t_wf * 1e6:10.0f} {by / t_wf / 1e3:7.0f} | {t_ref:13.1f} "
                  f"{B / t_ref * 1e6:10.0f} {by / t_ref / 1e3:7.0f} | {t_ref / t_wf:8.1f} {diff:10.2e}", flush=True)
            del pcm
            torch.cuda.empty_cache()
    # (iii) the same reference algorithm on the host cores (B in {1, 16, 64, 256})
    torch.set_num_threads(os.cpu_count())
    print(f"\nreference algorithm (torch.stft pipeline) on the host, fp32, {torch.get_num_threads()} threads")
    for n_mels in (80, 128):
        filt = mel_filters("cpu", n_mels)
        for B in (1, 16, 64, 256):
            pcm = torch.randn(B, 480000) * 0.1
            torch_stft_logmel(pcm[:1], filt)
            import time
            t0 = time.perf_counter()
            reps = 3 if B <= 16 else 1
            for _ in range(reps):
                torch_stft_logmel(pcm, filt)
            t = (time.perf_counter() - t0) / reps
            by = B * (480000 * 4 + n_mels * 3000 * 4)
            print(f"{B:6d} {n_mels:5d} | {t * 1e6:12.0f} us {B / t:10.0f} clips/s {by / t / 1e9:7.2f} GB/s", flush=True)


if __name__ == "__main__":
    main()
