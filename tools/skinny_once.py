"""A handful of decode-shape GEMM launches for an ncu capture (tools only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper-flamingo_b200"))
import torch
from whisper import _native as nv
m, n, k = 128, 1280, int(sys.argv[1]) if len(sys.argv) > 1 else 5120
ws = [torch.randn(n, k, device="cuda").bfloat16() * 0.02 for _ in range(8)]
a = torch.randn(m, k, device="cuda").bfloat16()
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for i in range(8):
    nv.linear(a, ws[i], out, tile_hint=32)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
