"""Data-parallel result equality on real GPUs: every rank decodes its own
contiguous block of clips on its own B200 and one NCCL all_gather returns the whole batch in clip order - the tokens
must equal the single-GPU decode of the same batch.  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped otherwise.
The world-size-2 gloo tests in test_parallel_cpu.py cover the same host logic without GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "whisper-flamingo_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import whisper
    from helpers import build_model
    from whisper._synthetic import synthetic_features, synthetic_pcm
    from whisper.parallel import decode_sharded, shard_range
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        model = build_model(gated=True, device=dev)
        n = 5  # ragged split: 3 | 2
        pcm = synthetic_pcm(n, seed=1234)
        mel = torch.stack([whisper.log_mel_spectrogram(p.to(dev)) for p in pcm])
        feat = synthetic_features(n, n_frames=60, dim=1024, seed=4321).to(dev)
        opt = whisper.DecodingOptions(language="en", without_timestamps=True, sample_len=10, fp16=False)
        toks, lps, nsp = decode_sharded(model, mel, opt, x_v=feat)
        lo, hi = shard_range(n, rank, world)
        local = whisper.decode(model, mel[lo:hi], opt, x_v=feat[lo:hi])
        ok_local = all(toks[lo + i, : len(r.tokens)].tolist() == r.tokens for i, r in enumerate(local))
        result = {"rank": rank, "tokens": toks.cpu(), "lps": lps, "nsp": nsp, "ok_local": ok_local}
        if rank == 0:
            full = whisper.decode(model, mel, opt, x_v=feat)
            result["full_tokens"] = [r.tokens for r in full]
            result["full_lps"] = [r.avg_logprob for r in full]
            result["full_nsp"] = [r.no_speech_prob for r in full]
        torch.save(result, os.path.join(out_dir, f"rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_decode_sharded_two_gpus_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    assert r0["ok_local"] and r1["ok_local"]
    assert torch.equal(r0["tokens"], r1["tokens"])                    # every rank receives the whole batch
    for i, want in enumerate(r0["full_tokens"]):                       # ... in clip order, equal to one GPU's decode
        assert r0["tokens"][i, : len(want)].tolist() == want
    assert max(abs(a - b) for a, b in zip(r0["lps"], r0["full_lps"])) < 1e-5
    assert max(abs(a - b) for a, b in zip(r0["nsp"], r0["full_nsp"])) < 1e-6
