"""the fork's teacher-forced calling mode (whisper-flamingo_*.py:256-266, trilingual.py:256,304):
    audio_features = model.encoder(mel);  logits = model.decoder(dec_input_ids, audio_features, xt_list=[f_1 .. f_n])
one batched pass over T_text tokens with several feature tensors, no autoregression.  The same call on the engine and on
the UNMODIFIED reference (torch eager, its half mode) on the same GPU, synthetic inputs, random-init weights.
usage: python tools/teacher_forced_bench.py [workload=medium] [batch=32] [t_text=64] [num_langs=3] [t_x=448]"""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisper-flamingo_b200")):
    sys.path.insert(0, p)
import torch
import whisper
from bench import FEAT_DIM, dims_for
from whisper._synthetic import init_synthetic_fast_

workload = sys.argv[1] if len(sys.argv) > 1 else "medium"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T_TEXT = int(sys.argv[3]) if len(sys.argv) > 3 else 64
LANGS = int(sys.argv[4]) if len(sys.argv) > 4 else 3
T_X = int(sys.argv[5]) if len(sys.argv) > 5 else 448
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dims = dims_for(workload)


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


g = torch.Generator(device=dev).manual_seed(7)
mel = torch.randn(B, dims["n_mels"], 3000, generator=g, device=dev)
feats = [torch.randn(B, T_X, FEAT_DIM, generator=g, device=dev) for _ in range(LANGS)]
ids = torch.randint(1000, 40000, (B, T_TEXT), generator=g, device=dev)

model = whisper.Whisper(whisper.ModelDimensions(**dims), 0.0, False, 256, 1, FEAT_DIM, LANGS).to(dev).eval()
init_synthetic_fast_(model, seed=0)


@torch.no_grad()
def ours():
    xa = model.encoder(mel.to(torch.bfloat16))
    return model.decoder(ids, xa, xt_list=feats)


@torch.no_grad()
def ours_decoder(xa):
    return model.decoder(ids, xa, xt_list=feats)


ms, logits = timeit(ours, 5)
xa = model.encoder(mel.to(torch.bfloat16))
ms_dec, _ = timeit(lambda: ours_decoder(xa), 10)
print(f"{workload} AV, {LANGS} feature tensors of {T_X} x {FEAT_DIM}, B={B}, {T_TEXT} teacher-forced tokens, bf16")
print(f"  engine:    encoder + decoder {ms:8.1f} ms = {30.0 * B / ms * 1e3:8.0f} audio-s/s | decoder pass alone "
      f"{ms_dec:7.1f} ms = {B * T_TEXT / ms_dec * 1e3:9.0f} tokens/s   (logits {tuple(logits.shape)} {logits.dtype})")
del model, logits
torch.cuda.empty_cache()

from baseline import reference_arm

ref = reference_arm.load()
if ref is None:
    print("  reference: not available on this box (baseline/_ref absent)")
    sys.exit(0)
from ref_whisper.model import ModelDimensions, Whisper

with contextlib.redirect_stdout(io.StringIO()):
    with torch.device(dev):
        rmodel = Whisper(ModelDimensions(**dims), 0.0, False, 256, 1, FEAT_DIM, LANGS).eval()
with torch.no_grad():
    for name, p in rmodel.named_parameters():
        if name.endswith("_gate"):
            p.fill_(0.5)
            p.data = p.data.half()      # the upstream notebook's recipe for half mode
        elif p.dim() >= 2:
            p.normal_(0, 0.02)
    rmodel.decoder.positional_embedding.normal_(0, 0.01)


@torch.no_grad()
def theirs():
    xa = rmodel.encoder(mel.half())
    return rmodel.decoder(ids, xa, xt_list=feats)


@torch.no_grad()
def theirs_decoder(xa):
    return rmodel.decoder(ids, xa, xt_list=feats)


rb = min(B, 16)                         # the reference materialises fp32 [B, H, 1500, 1500] scores per encoder layer
mel_r, feats_r, ids_r = mel[:rb], [f[:rb] for f in feats], ids[:rb]
mel, feats, ids = mel_r, feats_r, ids_r
ms_r, _ = timeit(theirs, 2)
xa_r = rmodel.encoder(mel.half())
ms_rd, _ = timeit(lambda: theirs_decoder(xa_r), 3)
print(f"  reference: encoder + decoder {ms_r:8.1f} ms = {30.0 * rb / ms_r * 1e3:8.0f} audio-s/s | decoder pass alone "
      f"{ms_rd:7.1f} ms = {rb * T_TEXT / ms_rd * 1e3:9.0f} tokens/s   (unmodified, torch eager, half mode, B={rb})")
