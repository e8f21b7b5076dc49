import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper-flamingo_b200"))
import torch
from whisper import _native as nv
B, H, T = 16, 12, 1500
d = H * 64
qkv = torch.randn(B * T, 3 * d, device="cuda").bfloat16()
out = torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    nv.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, T, T, H, False)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
