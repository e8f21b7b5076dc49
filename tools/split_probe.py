"""Decode-loop time vs number of concurrent sub-batches (WF_DECODE_SPLIT): large-v2 AV, B=128, 64 tokens."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "whisper-flamingo_b200"))
import torch
import bench
import whisper
from whisper import _engine
from whisper._synthetic import synthetic_features, synthetic_pcm

os.environ["WF_TIMING"] = "1"
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "large-v2"
B = bench.WORKLOADS[wl][3]
model = bench.build_model(wl, dev)
opt = bench.options()
pcm = synthetic_pcm(B, bench.N_SAMPLES, seed=1234).to(dev)
feat = synthetic_features(B, bench.T_X, bench.FEAT_DIM, seed=4321).to(dev)
ref = None
for split in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4").split(",")]:
    os.environ["WF_DECODE_SPLIT"] = str(split)
    _engine.clear_sessions()
    torch.cuda.empty_cache()
    for it in range(3):
        res = bench.hot_path_step(model, pcm, feat, opt)
        torch.cuda.synchronize()
    toks = [r.tokens for r in res]
    if ref is None:
        ref = toks
    same = sum(a == b for a, b in zip(ref, toks))
    print(f"split={split}: phases {dict((k, round(v, 1)) for k, v in _engine.PhaseTimer.last.items())} "
          f"rows identical to split=1: {same}/{len(toks)}", flush=True)
