"""A/B timing of the log-mel kernel chain (WF_LOGMEL_V1=1 selects the first layout). usage: python tools/mel_ab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-flamingo_b200"))
import torch
import whisper


def main():
    tag = "v1" if os.environ.get("WF_LOGMEL_V1") else ("v2" if os.environ.get("WF_LOGMEL_V2") else "v3")
    for n_mels in (80, 128):
        for B in (1, 4, 16, 128, 1024):
            pcm = torch.randn(B, 480000, device="cuda") * 0.1
            by = B * (480000 * 4 + n_mels * 3000 * 4)
            fn = lambda: whisper.log_mel_spectrogram(pcm, n_mels=n_mels, per_clip_max=True)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            iters = 50 if B <= 128 else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            print(f"{tag} mels={n_mels:3d} B={B:5d}  {us:9.1f} us  {B / us * 1e6:10.0f} clips/s  {by / us / 1e3:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
