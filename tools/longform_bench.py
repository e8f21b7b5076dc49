"""Long-form throughput: transcribe_batch (the current 30-s window of every recording as one batch per round) against
transcribe one recording at a time (what the reference does, transcribe.py:234-377).  Synthetic recordings of different
lengths, random-init weights (the text is noise; the window / segmentation / fallback machinery runs as it would).
usage: python tools/longform_bench.py [workload=small] [n_recordings=32] [minutes=3]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisper-flamingo_b200")):
    sys.path.insert(0, p)
import torch
import whisper
from bench import build_model, FEAT_DIM
from whisper._synthetic import synthetic_features, synthetic_pcm

workload = sys.argv[1] if len(sys.argv) > 1 else "small"
n_rec = int(sys.argv[2]) if len(sys.argv) > 2 else 32
minutes = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = build_model(workload, dev)
secs = [int(60 * minutes * (0.6 + 0.8 * i / max(1, n_rec - 1))) for i in range(n_rec)]       # 0.6 .. 1.4 x the mean
pcms = [synthetic_pcm(1, n_samples=s * 16000, seed=500 + i)[0].to(dev) for i, s in enumerate(secs)]
feats = [synthetic_features(1, n_frames=s * 25, dim=FEAT_DIM, seed=900 + i)[0].to(dev) for i, s in enumerate(secs)]
kw = dict(temperature=0.0, compression_ratio_threshold=None, logprob_threshold=None, no_speech_threshold=None,
          language="en", sample_len=64, verbose=None, condition_on_previous_text=False, suppress_tokens="-1")
total = float(sum(secs))


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


whisper.transcribe_batch(model, pcms[:2], x_v=feats[:2], **kw)      # warm-up: weight packing, first graphs
batch, t_b = timed(lambda: whisper.transcribe_batch(model, pcms, x_v=feats, **kw))
n_seq = min(4, n_rec)
seq, t_s = timed(lambda: [whisper.transcribe(model, pcms[i], x_v=feats[i], **kw) for i in range(n_seq)])
same = all(b["text"] == s["text"] for b, s in zip(batch[:n_seq], seq))
print(f"{workload} AV, {n_rec} recordings of {min(secs)}..{max(secs)} s ({total / 3600:.2f} h of audio), bf16, 64 tokens per window")
print(f"  transcribe_batch: {t_b:.2f} s  = {total / t_b:.0f} audio-s/s   ({sum(len(b['segments']) for b in batch)} segments)")
print(f"  transcribe, one recording at a time ({n_seq} of them): {t_s:.2f} s = {sum(secs[:n_seq]) / t_s:.0f} audio-s/s")
print(f"  texts of the first {n_seq} recordings identical between the two (bf16, batch-dependent rounding allowed): {same}")
