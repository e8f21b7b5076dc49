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

    last_local = {}

    def gather_tokens(results):
        toks = torch.tensor([r.tokens for r in results], dtype=torch.int32, device=dev)
        last_local["toks"] = toks
        return gather_token_matrix(toks, 50257)  # the only collective of the job: final result gather over NVLink

    def device_step():
        return gather_tokens(hot_path_step(model, pcm_dev, feat_dev, opt))

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_step():
        # host inputs -> tokens on the host, all through the public API.  The PCM goes up first on the compute stream;
        # the lip features (393 of the 639 MB) follow on a copy stream while the log-mel frontend and the encoder of
        # the SAME step run - the decoder is the first consumer of the features (whisper.decode takes encoded audio,
        # like the reference: whisper/decoding.py:663-676).
        import whisper
        main = torch.cuda.current_stream()
        p = pcm_host.to(dev, non_blocking=True)
        pcm_up = torch.cuda.Event()
        pcm_up.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(pcm_up)          # one H2D transfer at a time: the PCM is needed first
            f = feat_host.to(dev, non_blocking=True)
            feat_up = torch.cuda.Event()
            feat_up.record(copy_stream)
        mel = whisper.log_mel_spectrogram(p, n_mels=80, per_clip_max=True)
        audio = model.embed_audio(mel.to(torch.bfloat16))
        main.wait_event(feat_up)
        f.record_stream(main)
        toks = gather_tokens(whisper.decode(model, audio, opt, x_v=f))
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
    phases = None
    if os.environ.get("WF_TIMING", "0") == "1":   # of the device-resident run (the e2e step feeds decode() encoded audio)
        from whisper._engine import PhaseTimer
        phases = {k: round(v, 2) for k, v in PhaseTimer.last.items()}
    e2e_step()
    ms_e2e, toks_e2e = timed(e2e_step, args.steps)
    assert toks.shape == (B * world, SAMPLE_LEN), toks.shape
    assert torch.equal(toks.cpu(), toks_e2e), "device-resident and end-to-end runs decoded different tokens"
    # data-parallel result check: the gathered matrix holds rank r's own tokens in rows [r * B, (r + 1) * B)
    assert torch.equal(toks_e2e[rank * B:(rank + 1) * B], last_local["toks"].cpu()), \
        f"rank {rank}: gathered rows differ from the locally decoded tokens"

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
            eager = None
            try:
                eager = gpu_eager_reference(args.workload, dev)
            except Exception as e:  # noqa: BLE001 - a comparator must never take the product line down
                eager = {"error": repr(e)[:200]}
            if eager is not None:
                line["gpu_eager_reference"] = eager
            line["cpu_baseline"] = cpu_baseline(args.workload, model, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# DRAM bytes (read + write) per launch of a kernel family, from the committed `ncu --set full` capture of this same
# configuration (profiles/r02_ncu_full_latent_pair.txt: large-v2 AV, B=128, greedy): the cross-attention launch over 1500
# encoder rows reads 498.5 MB + writes 5.6 MB against 491.5 MB of source rows, the x-attention launch over 750 feature
# rows 252.8 + 5.5 MB against 245.8 MB - every source row crosses the HBM interface once (the round-1 two-pass kernel:
# 612.7 MB).  Reported as `traffic` with `traffic_source: "profile"`: ncu cannot run inside the timed region.
NCU_TRAFFIC = {"latent_attention_T1500": 504.1e6, "latent_attention_T750": 258.3e6}


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
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture - only for the configuration
    # that was captured
    traffic = None
    if B == 128 and not BEAM and len(model.decoder.blocks) == 32:
        traffic = NCU_TRAFFIC.get(top["kernel"])
    roof = {"kernel": top["kernel"], "bound": "tensor" if tensor_bound else "hbm", "achieved": round(ach, 1),
            "peak": peak, "unit": unit, "frac": round(ach / peak, 4), "traffic": traffic,
            "traffic_source": None if traffic is None else "profile (profiles/r02_ncu_full_latent_pair.txt)",
            "algorithmic_per_launch": round((f["flops"] if tensor_bound else f["bytes"]) / f["launches"], 1),
            "avg_launch_ms": round(f["ms"] / f["launches"], 4), "peak_source": pk["source"],
            "share_of_step": top["share_above_launch_floor"], "launch_floor_us": round(floor_ms * 1e3, 2)}
    # An event pair around ONE eager launch of a 100-us kernel also holds ~10 us of launch latency (the GPU idles until
    # the launch arrives: the same ~13 us the bracket around the 2.5-us step_advance kernel reads, `launch_floor_us`), and
    # the kernel starts cold, without the early loads programmatic dependent launch gives it in the replayed graph of the
    # timed region.  The dominant kernel is therefore timed again the way the timed region runs it - inside a CUDA graph -
    # and THAT is `achieved`; the eager figure stays next to it.
    in_graph = dominant_in_graph(model, top["kernel"], pk)
    if in_graph:
        roof["eager_bracketed"] = {"avg_launch_ms": roof["avg_launch_ms"], "achieved": roof["achieved"], "frac": roof["frac"]}
        roof["avg_launch_ms"], roof["achieved"], roof["frac"] = in_graph["avg_launch_ms"], in_graph["achieved"], in_graph["frac"]
        roof["how"] = in_graph["how"]
    return {"roofline": roof, "kernels": table, "profiled_step_ms": round(total, 2)}


def dominant_in_graph(model, family, pk, reps=32):
    """The dominant kernel as the timed region runs it: inside a CUDA graph (programmatic dependent launch lets its
    first loads start under the previous kernel), 32 back-to-back launches on the decode session's own buffers
    (491.5 MB of source rows per launch: larger than L2), bracketed by CUDA events.  The eager, per-launch-bracketed
    figure above starts every launch cold.  Only for the latent attention families."""
    from whisper import _engine, _native as nv
    if not family.startswith("latent_attention_T"):
        return None
    sess = _engine.last_session(model.decoder)
    if sess is None or not getattr(sess, "latent", False):
        return None
    t = int(family.rsplit("T", 1)[1])
    src = sess.xa_src if sess.xa_src.shape[1] == t else next((x for x in getattr(sess, "x_src", []) if x.shape[1] == t), None)
    if src is None:
        return None
    H = sess.p.n_head
    ctx = sess.ctx if sess.split else sess.ctx[0]

    def launch():
        nv.latent_attention(sess.qp, src, ctx, H, ml=sess.ctx_ml if sess.split else None)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        launch()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            launch()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ach = src.numel() * 2 / (ms * 1e-3) / 1e9
    return {"avg_launch_ms": round(ms, 4), "achieved": round(ach, 1), "frac": round(ach / pk["hbm"], 4), "launches": reps,
            "how": "CUDA graph of back-to-back launches on the session's buffers, CUDA events around the replay"}


# ----------------------------------------------------------------------------- reference arm / CPU baseline
# Preferred: the UNMODIFIED reference package (baseline/_ref staged copy, or /root/reference) driven by
# baseline/reference_arm.py -> kind "reference".  Only when neither exists on the box: the oracle port of the same
# algorithm (oracle/, pinned to the reference by oracle/pin_reference.py) -> kind "port".
# The reference keeps no KV cache: every decode step recomputes the decoder over ALL tokens and re-projects the cross /
# x-attention K,V of all 1500 + 750 source frames (decoding.py:155-164), so one 64-token clip of large-v2 costs ~30 s on
# the host cores.  A bench step therefore times a BOUNDED SAMPLE of one clip: log-mel + encoder in full, and the decode
# step at sampled-token positions spread over the whole range (first / middle / last - the cost is affine in the
# position, so their mean is the mean over all 64); `value` extrapolates to the full 64-step clip, `ms_per_step` is the
# wall time of the sample itself.
REF_POSITIONS = (0, SAMPLE_LEN // 2, SAMPLE_LEN - 1)


def _oracle_spec(model, sample_len):
    from oracle import decode as odec
    from whisper.decoding import DecodingTask
    task = DecodingTask(model, options())
    tk = task.tokenizer
    return odec.DecodeSpec(initial_tokens=tuple(task.initial_tokens), eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech,
                           suppress_tokens=tuple(task._get_suppress_tokens()),
                           blank_tokens=tuple(tk.encode(" ") + [tk.eot]), sample_len=sample_len,
                           n_ctx=model.dims.n_text_ctx)


class CpuArm:
    """One clip of the workload on the host cores, fp32, all threads."""

    def __init__(self, workload):
        from whisper._synthetic import synthetic_features, synthetic_pcm
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.workload = workload
        self.pcm = synthetic_pcm(1, N_SAMPLES, seed=1234)
        self.feat = synthetic_features(1, T_X, FEAT_DIM, seed=4321)
        from baseline import reference_arm as ra
        self.ra = ra
        self.ref = ra.load()
        if self.ref is not None:
            self.kind = "reference"
            self.where = ra.locate()
            _, self.model = ra.build_model(dims_for(workload), "cpu")
        else:
            self.kind = "port"
            self.where = "oracle/ (reference package not present on this box)"
            import whisper
            from oracle import model as om
            m = whisper.Whisper(whisper.ModelDimensions(**dims_for(workload)), 0.0, False, 256, 1, FEAT_DIM, 1).eval()
            with torch.no_grad():
                for name, p in m.named_parameters():
                    if name.endswith("_gate"):
                        p.fill_(0.5)
                    elif p.dim() >= 2:
                        p.normal_(0, 0.02)
                m.decoder.positional_embedding.normal_(0, 0.01)
            self.model = m
            self.sd = om.cast_state_dict_fp32(m.state_dict())
            self.dims = om.Dims(**dims_for(workload))

    def sample(self, positions):
        """-> dict(mel_s, enc_s, mean_step_s, full_s, wall_s)"""
        t0 = time.perf_counter()
        if self.kind == "reference":
            r = self.ra.time_hot_path(self.ref, self.model, self.pcm, self.feat, SAMPLE_LEN, positions)
        else:
            r = self._port_sample(positions)
        r["wall_s"] = time.perf_counter() - t0
        return r

    def _port_sample(self, positions):
        import numpy as np
        from oracle import decode as odec
        from oracle import mel as omel
        from oracle import model as om
        t0 = time.perf_counter()
        mel = torch.from_numpy(omel.log_mel_spectrogram(self.pcm.numpy(), 80, dtype=np.float32))
        t1 = time.perf_counter()
        with torch.no_grad():
            xa = om.encoder_forward(self.sd, self.dims, mel)
        t2 = time.perf_counter()
        spec = _oracle_spec(self.model, SAMPLE_LEN)
        g = torch.Generator().manual_seed(0)
        step_s = {}
        for pos in positions:
            tokens = torch.cat([torch.tensor([list(spec.initial_tokens)]),
                                torch.randint(1000, 40000, (1, pos), generator=g)], dim=1)
            s0 = time.perf_counter()
            with torch.no_grad():
                logits = om.decoder_forward(self.sd, self.dims, tokens, xa, xt_list=[self.feat])[:, -1]
                odec.apply_filters(spec, logits, tokens)
                odec.greedy_update(spec, tokens, logits, torch.zeros(1))
            step_s[pos] = time.perf_counter() - s0
        mean_step = sum(step_s.values()) / len(step_s)
        return {"mel_s": t1 - t0, "enc_s": t2 - t1, "step_s": step_s, "mean_step_s": mean_step,
                "full_s": (t1 - t0) + (t2 - t1) + SAMPLE_LEN * mean_step}

    def describe(self, r, positions):
        what = ("the UNMODIFIED reference package (%s)" % self.where if self.kind == "reference"
                else "the oracle port of the reference algorithm (%s)" % self.where)
        return (f"1 clip of the same workload on {self.cores} host threads, fp32, {what}: log-mel ({r['mel_s']:.2f} s) "
                f"+ encoder ({r['enc_s']:.2f} s) in full, the no-KV-cache decode step timed at sampled-token positions "
                f"{list(positions)} of {SAMPLE_LEN} ({r['mean_step_s']:.3f} s mean; the cost is affine in the position) "
                f"and extrapolated to {SAMPLE_LEN} steps = {r['full_s']:.1f} s per clip; the sample itself took "
                f"{r['wall_s']:.1f} s")


def cpu_baseline(workload, model=None, budget_s=25.0):
    arm = CpuArm(workload)
    r = arm.sample(REF_POSITIONS[1:2])          # warm-up (thread pools, first-touch of 8.7 GB of weights)
    positions = REF_POSITIONS if r["wall_s"] * 2.2 <= budget_s else REF_POSITIONS[1:2]
    r = arm.sample(positions)
    return {"value": 30.0 / r["full_s"], "unit": "audio-s/s", "cores": arm.cores, "kind": arm.kind,
            "sample": arm.describe(r, positions)}


def gpu_eager_reference(workload, dev, batch=16):
    """The unmodified reference on the SAME GPU (torch eager, its own half mode: fp32 master weights cast per call,
    fp16 activations, no KV cache, no fused kernels) - the honest GPU comparator of SURVEY.md section 8d.  Bounded:
    `batch` clips, decode steps at REF_POSITIONS, extrapolated like the CPU arm.  None when the package is absent."""
    from baseline import reference_arm as ra
    from whisper._synthetic import synthetic_features, synthetic_pcm
    if ra.load() is None:
        return None
    ref, model = ra.build_model(dims_for(workload), dev, half=True)
    pcm = synthetic_pcm(batch, N_SAMPLES, seed=1234).to(dev)
    feat = synthetic_features(batch, T_X, FEAT_DIM, seed=4321).to(dev)
    sync = lambda: torch.cuda.synchronize(dev)  # noqa: E731
    ra.time_hot_path(ref, model, pcm, feat, SAMPLE_LEN, REF_POSITIONS[1:2], half=True, sync=sync)  # warm-up
    r = ra.time_hot_path(ref, model, pcm, feat, SAMPLE_LEN, REF_POSITIONS, half=True, sync=sync)
    del model
    torch.cuda.empty_cache()
    return {"value": 30.0 * batch / r["full_s"], "unit": "audio-s/s", "batch": batch, "dtype": "fp16 activations, "
            "fp32 master weights cast per call (the reference's half mode)",
            "sample": f"{batch} clips, torch eager on the same B200: log-mel {r['mel_s'] * 1e3:.1f} ms + encoder "
                      f"{r['enc_s'] * 1e3:.1f} ms in full, no-KV-cache decode step at positions {list(REF_POSITIONS)} "
                      f"({r['mean_step_s'] * 1e3:.1f} ms mean) extrapolated to {SAMPLE_LEN} steps = {r['full_s']:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.workload)
    warm = max(args.warmup, 1)
    r = arm.sample(REF_POSITIONS[1:2])
    # K timed steps must end within a few minutes: three decode positions per step when they fit, else the midpoint only
    positions = REF_POSITIONS if r["wall_s"] * 2.2 * args.steps <= 170.0 else REF_POSITIONS[1:2]
    for _ in range(warm - 1):
        arm.sample(positions)
    t0 = time.perf_counter()
    samples = [arm.sample(positions) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    full = sum(x["full_s"] for x in samples) / len(samples)
    value = 30.0 / full
    B = args.batch or WORKLOADS[args.workload][3]
    last = samples[-1]
    print(json.dumps({
        "impl": "reference", "metric": "audio-sec/sec, Whisper-Flamingo AV greedy decode (mel->tokens)",
        "value": value, "unit": "audio-s/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": warm, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} AV (gated x-attn, n_text_ctx=768), B={B}/GPU, 30 s clips, "
                               f"750x1024 features, greedy {SAMPLE_LEN} tokens (EOT suppressed)",
                   "note": "CPU arm: one clip per step (throughput per clip is independent of B); each step is a "
                           "bounded sample, `value` = 30 s / (time of one full clip extrapolated from the sample), "
                           "`ms_per_step` = wall time of the sample",
                   "extrapolated_s_per_clip": full, "timed_decode_positions": list(positions)},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.describe(last, positions)},
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
