"""cProfile of a warm transcribe_batch call (large-v2 AV, 64 recordings of 72..168 s): where the host spends its time.
Measured on the final tree: 4.3 s for 288 windows in 6 rounds, of which 3.9 s are waits on the GPU (decode loops 2.4 s,
stream-ordered copies behind the encoder 1.2 s, result read-back 0.3 s): the driver is GPU-bound; rounds of 64 windows cost
~65 % of a 128-window bench step.  usage: python tools/longform_profile.py"""
import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisper-flamingo_b200")):
    sys.path.insert(0, p)
import torch, whisper
from bench import build_model, FEAT_DIM
from whisper._synthetic import synthetic_features, synthetic_pcm
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
model = build_model("large-v2", dev)
n_rec, minutes = 64, 2.0
secs = [int(60 * minutes * (0.6 + 0.8 * i / max(1, n_rec - 1))) for i in range(n_rec)]
pcms = [synthetic_pcm(1, n_samples=s * 16000, seed=500 + i)[0].to(dev) for i, s in enumerate(secs)]
feats = [synthetic_features(1, n_frames=s * 25, dim=FEAT_DIM, seed=900 + i)[0].to(dev) for i, s in enumerate(secs)]
kw = dict(temperature=0.0, compression_ratio_threshold=None, logprob_threshold=None, no_speech_threshold=None,
          language="en", sample_len=64, verbose=None, condition_on_previous_text=False, suppress_tokens="-1")
whisper.transcribe_batch(model, pcms[:2], x_v=feats[:2], **kw)
whisper.transcribe_batch(model, pcms, x_v=feats, **kw)   # warm: sessions exist
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
whisper.transcribe_batch(model, pcms, x_v=feats, **kw)
torch.cuda.synchronize(); pr.disable(); print("second call", time.perf_counter() - t0, "s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
