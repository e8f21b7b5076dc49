"""One log-mel call at B clips (for ncu). usage: python tools/mel_once.py [B] [n_mels]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-flamingo_b200"))
import torch
import whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_mels = int(sys.argv[2]) if len(sys.argv) > 2 else 80
pcm = torch.randn(B, 480000, device="cuda") * 0.1
for _ in range(3):
    out = whisper.log_mel_spectrogram(pcm, n_mels=n_mels, per_clip_max=True)
torch.cuda.synchronize()
print(float(out.mean()))
