#!/usr/bin/env python
"""Benchmark of the Whisper-Flamingo AV inference hot path on B200 (driver contract: see task brief).

A "step" = one pass of the hot path over one batch of synthetic 30-s clips:
    PCM -> log-mel -> AudioEncoder -> cross / x-attn K,V precompute -> KV-cached greedy decode -> token ids.
metric = audio-seconds transcribed per second (BASELINE.json), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload large-v2|medium|small|tiny] [--batch B]
  python bench.py --impl reference ...   # the reference algorithm (CPU oracle port) on the host cores

N > 1: launched by torchrun, one rank per GPU, each rank decodes its own batch (weak scaling, no collective
in the layer loop); the only NCCL traffic is the final token gather.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "whisper-flamingo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (d, heads, layers, default per-GPU batch); n_text_ctx = 768 so that 750 feature frames fit
    "tiny": (384, 6, 4, 16),
    "small": (768, 12, 12, 16),
    "medium": (1024, 16, 24, 64),
    "large-v2": (1280, 20, 32, 128),
}
SAMPLE_LEN = 64
T_X, FEAT_DIM = 750, 1024
N_SAMPLES = 480000


def dims_for(name):
    d, h, l, _ = WORKLOADS[name]
    return dict(n_mels=80, n_audio_ctx=1500, n_audio_state=d, n_audio_head=h, n_audio_layer=l, n_vocab=51865,
                n_text_ctx=768, n_text_state=d, n_text_head=h, n_text_layer=l)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, read in-process through NVML every 250 ms
    (an external `nvidia-smi -lms 200` loop was measured to perturb the timed region by up to 30 %)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period_s: float = 0.25):
        self.sm, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.period = period_s
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while not self._stop.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for bit, n in self.REASONS.items() if self.mask & bit), "samples": len(self.sm)}


# ----------------------------------------------------------------------------- product arm
def build_model(workload: str, device):
    import whisper
    from whisper._synthetic import init_synthetic_fast_
    model = whisper.Whisper(whisper.ModelDimensions(**dims_for(workload)), 0.0, False, 256, 1, FEAT_DIM, 1)
    model = model.to(device).eval()
    init_synthetic_fast_(model, seed=0)
    return model


BEAM = 0  # --beam N: beam search instead of greedy (BASELINE config 3: medium, beam 5)


def options():
    import whisper
    # EOT is suppressed so that every clip costs exactly SAMPLE_LEN steps
    return whisper.DecodingOptions(language="en", task="transcribe", without_timestamps=True, temperature=0.0,
                                   sample_len=SAMPLE_LEN, suppress_tokens="-1,50257", suppress_blank=True, fp16=True,
                                   beam_size=BEAM or None)


def hot_path_step(model, pcm_dev, feat_dev, opt):
    """Device-resident step through the public API; returns the per-clip token lists' device tensor stand-in."""
    import whisper
    mel = whisper.log_mel_spectrogram(pcm_dev, n_mels=80, per_clip_max=True)
    return whisper.decode(model, mel, opt, x_v=feat_dev)


def run_product(args):
    import whisper  # noqa: F401
    from whisper import _native as nv
    from whisper._synthetic import synthetic_features, synthetic_pcm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or WORKLOADS[args.workload][3]
    model = build_model(args.workload, dev)
    opt = options()
    pcm_host = synthetic_pcm(B, N_SAMPLES, seed=1234 + rank).pin_memory()
    feat_host = synthetic_features(B, T_X, FEAT_DIM, seed=4321 + rank).pin_memory()
    pcm_dev, feat_dev = pcm_host.to(dev), feat_host.to(dev)

    if args.single_step:
        hot_path_step(model, pcm_dev, feat_dev, opt)  # not profiled: one-time weight packing, graph capture
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the hot path is captured, not the weight init
        res = hot_path_step(model, pcm_dev, feat_dev, opt)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"single_step": True, "workload": args.workload, "batch": B, "tokens0": res[0].tokens[:8],
                          "launches": nv.kernel_launch_count()}), flush=True)
        return

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from whisper.parallel import gather_token_matrix

    def gather_tokens(results):
        toks = torch.tensor([r.tokens for r in results], dtype=torch.int32, device=dev)
        return gather_token_matrix(toks, 50257)  # the only collective of the job: final result gather over NVLink

    def device_step():
        return gather_tokens(hot_path_step(model, pcm_dev, feat_dev, opt))

    def e2e_step():
        p = pcm_host.to(dev, non_blocking=True)
        f = feat_host.to(dev, non_blocking=True)
        toks = gather_tokens(hot_path_step(model, p, f, opt))
        return toks.cpu()  # device -> host read of the step's result

    def timed(fn, steps):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            out = fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(max(args.warmup, 3)):
        device_step()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = nv.kernel_launch_count()
    ms, toks = timed(device_step, args.steps)
    launches = nv.kernel_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    e2e_step()
    ms_e2e, toks_e2e = timed(e2e_step, args.steps)
    phases = None
    if os.environ.get("WF_TIMING", "0") == "1":
        from whisper._engine import PhaseTimer
        phases = {k: round(v, 2) for k, v in PhaseTimer.last.items()}
    assert toks.shape == (B * world, SAMPLE_LEN), toks.shape
    assert torch.equal(toks.cpu(), toks_e2e), "device-resident and end-to-end runs decoded different tokens"

    audio_s = 30.0 * B * world * args.steps
    line = {
        "metric": "audio-sec/sec, Whisper-Flamingo AV greedy decode (mel->tokens)", "value": audio_s / (ms / 1e3),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload} AV (gated x-attn, n_text_ctx=768), B={B}/GPU, 30 s clips, "
                               f"750x1024 features, {'beam-%d' % BEAM if BEAM else 'greedy'} {SAMPLE_LEN} tokens "
                               f"(EOT suppressed), bf16",
                   "global_batch": B * world, "parallelism": f"dp{world}", "l2": "inputs_larger_than_l2",
                   "random_init_weights": True},
        "e2e": {"value": audio_s / (ms_e2e / 1e3), "unit": "audio-s/s",
                "h2d_bytes_per_step": (pcm_host.numel() * 4 + feat_host.numel() * 4) * world,
                "d2h_bytes_per_step": B * world * SAMPLE_LEN * 4},
        "gpu_launches": launches, "clocks": clocks,
    }
    if phases:
        line["phases_ms"] = phases
    if rank == 0:
        if not args.no_profile:
            line.update(profile_kernels(model, pcm_dev, feat_dev, opt, B, ms / args.steps))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, model, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# DRAM bytes (read + write) per launch of a kernel family from the committed `ncu --set full` captures under profiles/
# (large-v2 AV, B=128, greedy): attention_decode alternates a cross-attention (988.0 MB) and an x-attention (499.2 MB)
# launch with the K/V-cache path and is x-attention only (497.0 MB) with the latent path; latent_attention: 608.7 MB read
# + 4.0 MB written against 491.5 MB of source rows - part of the second pass misses L2 (r01_ncu_full_latent_attn.txt)
NCU_TRAFFIC = {"attention_decode": 497.0e6, "latent_attention": 612.7e6}


def launch_floor_ms(dev, n=256):
    """Average event-bracketed time of one eager launch of the smallest libwf kernel (step_advance, one thread block)."""
    from whisper import _native as nv
    state = torch.zeros(8, dtype=torch.int32, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        e0.record()
        nv.step_advance(state, 1)
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in ev) / n


def profile_kernels(model, pcm_dev, feat_dev, opt, B, step_ms):
    """One extra instrumented step (every libwf launch bracketed by CUDA events on its stream, CUDA graph off)
    -> per-kernel-family time share and achieved rate; `roofline` describes the dominant family."""
    from whisper import _native as nv
    pk = peaks()
    rec = []
    nv.set_profiler(lambda fam, work, e0, e1: rec.append((fam, work, e0, e1)))
    os.environ["WF_NO_GRAPH"] = "1"          # every decode-step kernel becomes an individually timed launch
    os.environ["WF_NO_SESSION_CACHE"] = "1"  # ... in a fresh session (the cached one replays its captured graph)
    try:
        hot_path_step(model, pcm_dev, feat_dev, opt)
        torch.cuda.synchronize()
    finally:
        nv.set_profiler(None)
        os.environ["WF_NO_GRAPH"] = "0"
        os.environ["WF_NO_SESSION_CACHE"] = "0"
    fams = {}
    for fam, work, e0, e1 in rec:
        f = fams.setdefault(fam, {"ms": 0.0, "launches": 0, "flops": 0, "bytes": 0})
        f["ms"] += e0.elapsed_time(e1)
        f["launches"] += 1
        f["flops"] += work.get("flops", 0)
        f["bytes"] += work.get("bytes", 0)
    total = sum(f["ms"] for f in fams.values()) or 1.0
    # An event pair around ONE eagerly launched kernel also times the launch gap: a floor of a few microseconds that the
    # replayed CUDA graph of the real step does not pay and that inflates the families made of thousands of tiny
    # launches.  The floor is measured here by bracketing the 2.5-us `step_advance` kernel the same way (~13 us: the
    # Python / ctypes / event-record cost per launch, during which the GPU idles).  A bracket reads about
    # max(kernel time, floor), so the time above the floor is a LOWER bound of a family's in-graph time; families are
    # RANKED by it (`share_above_launch_floor`) so that 21 500 eager launches of 8-us GEMMs do not outrank the kernel
    # that leads the ncu launch list under profiles/; every reported rate still uses the raw event time.
    floor_ms = launch_floor_ms(pcm_dev.device)
    for f in fams.values():
        f["ms_corr"] = max(f["ms"] - f["launches"] * floor_ms, 0.0) if f["launches"] >= 8 else f["ms"]
    total_corr = sum(f["ms_corr"] for f in fams.values()) or 1.0
    table = []
    for fam, f in sorted(fams.items(), key=lambda kv: -kv[1]["ms_corr"]):
        row = {"kernel": fam, "launches": f["launches"], "ms": round(f["ms"], 3), "share": round(f["ms"] / total, 4),
               "share_above_launch_floor": round(f["ms_corr"] / total_corr, 4)}
        if f["flops"]:
            row["tflops"] = round(f["flops"] / (f["ms"] * 1e-3) / 1e12, 1)
        if f["bytes"]:
            row["gbs"] = round(f["bytes"] / (f["ms"] * 1e-3) / 1e9, 1)
        table.append(row)
    top = table[0]
    tensor_bound = top["kernel"] in ("gemm_tc_bf16", "attention_full")
    f = fams[top["kernel"]]
    if tensor_bound:
        ach, peak, unit = f["flops"] / (f["ms"] * 1e-3) / 1e12, pk["tc_sustained"], "TFLOP/s"
    else:
        ach, peak, unit = f["bytes"] / (f["ms"] * 1e-3) / 1e9, pk["hbm"], "GB/s"
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/
    # r01_ncu_full_attn_decode_hm.txt: 988.0 MB for a cross-attention launch, 499.2 MB for an x-attention launch; the
    # family alternates the two, algorithmic 983.0 / 491.5 MB) - only for the configuration that was captured
    traffic = None
    if B == 128 and not BEAM and len(model.decoder.blocks) == 32:
        traffic = NCU_TRAFFIC.get(top["kernel"])
    roof = {"kernel": top["kernel"], "bound": "tensor" if tensor_bound else "hbm", "achieved": round(ach, 1),
            "peak": peak, "unit": unit, "frac": round(ach / peak, 4), "traffic": traffic,
            "algorithmic_per_launch": round((f["flops"] if tensor_bound else f["bytes"]) / f["launches"], 1),
            "avg_launch_ms": round(f["ms"] / f["launches"], 4), "peak_source": pk["source"],
            "share_of_step": top["share_above_launch_floor"], "launch_floor_us": round(floor_ms * 1e3, 2)}
    return {"roofline": roof, "kernels": table, "profiled_step_ms": round(total, 2)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _oracle_spec(model, sample_len):
    from oracle import decode as odec
    from whisper.decoding import DecodingTask
    task = DecodingTask(model, options())
    tk = task.tokenizer
    return odec.DecodeSpec(initial_tokens=tuple(task.initial_tokens), eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech,
                           suppress_tokens=tuple(task._get_suppress_tokens()),
                           blank_tokens=tuple(tk.encode(" ") + [tk.eot]), sample_len=sample_len,
                           n_ctx=model.dims.n_text_ctx)


def cpu_sample(workload, sd, dims, spec_fn, n_steps):
    """Reference algorithm on the host cores for ONE clip: log-mel + encoder in full, `n_steps` of the
    no-KV-cache decode loop; returns (mel_s, enc_s, per_step_s)."""
    from oracle import decode as odec
    from oracle import mel as omel
    from oracle import model as om
    from whisper._synthetic import synthetic_features, synthetic_pcm
    pcm = synthetic_pcm(1, N_SAMPLES, seed=1234).numpy()
    feat = synthetic_features(1, T_X, FEAT_DIM, seed=4321)
    t0 = time.perf_counter()
    import numpy as np
    mel = torch.from_numpy(omel.log_mel_spectrogram(pcm, 80, dtype=np.float32))
    t1 = time.perf_counter()
    with torch.no_grad():
        xa = om.encoder_forward(sd, dims, mel)
    t2 = time.perf_counter()
    spec = spec_fn(n_steps)
    with torch.no_grad():
        odec._run_group(sd, dims, spec, xa, feat, 1)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, (t3 - t2) / n_steps


def cpu_baseline(workload, model, budget_s=25.0, n_steps=None):
    from oracle import model as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = om.cast_state_dict_fp32(model.state_dict())
    dims = om.Dims(**dims_for(workload))
    if n_steps is None:
        n_steps = {"tiny": 8, "small": 4, "medium": 2, "large-v2": 2}[workload]
    mel_s, enc_s, step_s = cpu_sample(workload, sd, dims, lambda n: _oracle_spec(model, n), n_steps)
    full = mel_s + enc_s + SAMPLE_LEN * step_s
    return {"value": 30.0 / full, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 clip of the same workload, fp32, oracle port of the reference algorithm: log-mel "
                      f"({mel_s:.2f} s) + encoder ({enc_s:.2f} s) timed in full, {n_steps} of {SAMPLE_LEN} "
                      f"no-KV-cache decode steps timed ({step_s:.2f} s/step) and extrapolated linearly"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import whisper
    from whisper._synthetic import init_synthetic_
    from oracle import model as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = whisper.Whisper(whisper.ModelDimensions(**dims_for(args.workload)), 0.0, False, 256, 1, FEAT_DIM, 1).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():  # cheap CPU init (values do not affect the timing)
            if name.endswith("_gate"):
                p.fill_(0.5)
            elif p.dim() >= 2:
                p.normal_(0, 0.02)
        model.decoder.positional_embedding.normal_(0, 0.01)
    sd = om.cast_state_dict_fp32(model.state_dict())
    dims = om.Dims(**dims_for(args.workload))
    n_steps = {"tiny": 8, "small": 4, "medium": 2, "large-v2": 1}[args.workload]
    spec_fn = lambda n: _oracle_spec(model, n)
    times = []
    for i in range(max(args.warmup, 1) + args.steps):
        mel_s, enc_s, step_s = cpu_sample(args.workload, sd, dims, spec_fn, n_steps)
        if i >= max(args.warmup, 1):
            times.append(mel_s + enc_s + SAMPLE_LEN * step_s)
    full = sum(times) / len(times)
    value = 30.0 / full
    B = args.batch or WORKLOADS[args.workload][3]
    sample = (f"each step = 1 clip on the host cores (fp32, reference algorithm incl. its per-step full recompute): "
              f"log-mel + encoder in full, {n_steps} of {SAMPLE_LEN} decode steps timed, extrapolated linearly")
    print(json.dumps({
        "impl": "reference", "metric": "audio-sec/sec, Whisper-Flamingo AV greedy decode (mel->tokens)",
        "value": value, "unit": "audio-s/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} AV (gated x-attn, n_text_ctx=768), B={B}/GPU, 30 s clips, "
                               f"750x1024 features, greedy {SAMPLE_LEN} tokens (EOT suppressed)",
                   "note": "CPU arm: throughput is per clip, independent of B"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="large-v2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--beam", type=int, default=0, help="beam size (0 = greedy, the headline configuration)")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--single-step", action="store_true",
                    help="build the model, run ONE device-resident step and exit (the command profiled under ncu)")
    args = ap.parse_args()
    global BEAM
    BEAM = args.beam
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
