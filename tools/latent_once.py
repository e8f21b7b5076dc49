"""One launch of the latent cross-attention kernel at the headline shape (for ncu / -DLA_TIMING builds)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-flamingo_b200"))
import torch
from whisper import _native as nv

B, H, T = int(os.environ.get("B", 128)), int(os.environ.get("H", 20)), int(os.environ.get("T", 1500))
d = 64 * H
src = torch.randn(B, T, d, device="cuda").bfloat16()
qp = (torch.randn(B, H, d, device="cuda") * 0.05).bfloat16()
split = os.environ.get("SPLIT", "1") == "1" and nv.latent_split_supported(H)
ctx = torch.zeros(2 if split else 1, B, H, d, device="cuda", dtype=torch.bfloat16)
ml = torch.zeros(2, B, 32, 2, device="cuda") if split else None
for _ in range(3):
    nv.latent_attention(qp, src, ctx if split else ctx[0], H, ml=ml)
torch.cuda.synchronize()
print("ok", float(ctx.float().abs().mean()))
