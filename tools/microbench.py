"""Isolated kernel timings on the GPU (CUDA events, rotating buffers so weights / KV come from HBM, not L2).
usage: python tools/microbench.py gemm|attn|ln|mel|fa"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-flamingo_b200"))
import torch
from whisper import _native as nv


def timeit(fn, n_rot, iters=20):
    for i in range(3):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def graph_timeit(fn, n_rot, reps=64, iters=5):
    """Per-launch time (us) of fn inside a replayed CUDA graph of `reps` launches (no CPU launch cost)."""
    fn(0)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % n_rot)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3


def skinny():
    print("decode-step kernels inside a CUDA graph (true GPU time per launch, us)")
    m = 128
    x = torch.randn(m, 1280, device="cuda").bfloat16()
    w_ln, b_ln = torch.ones(1280, device="cuda"), torch.zeros(1280, device="cuda")
    xn = torch.empty_like(x)
    print(f"layernorm 128x1280 alone: {graph_timeit(lambda i: nv.layernorm(x, w_ln, b_ln, xn), 1):.2f} us")
    for (n, k) in ((1280, 1280), (2560, 1280), (5120, 1280), (1280, 5120), (51865, 1280)):
        n_rot = max(2, int(300e6 // (n * k * 2)) + 1)
        ws = [torch.randn(n, k, device="cuda").bfloat16() * 0.02 for _ in range(n_rot)]
        a = torch.randn(m, k, device="cuda").bfloat16()
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        row = f"M={m} N={n:6d} K={k:5d}: "
        for hint in (32, 64, 128):
            us = graph_timeit(lambda i: nv.linear(a, ws[i], out, tile_hint=hint), n_rot)
            row += f" bn{hint}: {us:6.2f}us {n * k * 2 / us / 1e3:6.0f}GB/s |"
        skw = torch.zeros(4096 + 8 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        for hint in (0, 32, 64):
            us = graph_timeit(lambda i: nv.linear(a, ws[i], out, tile_hint=hint, ws=skw), n_rot)
            row += f" splitK bn{hint}: {us:6.2f}us {n * k * 2 / us / 1e3:6.0f}GB/s |"
        if k == 1280:
            def pair(i):
                nv.layernorm(x, w_ln, b_ln, xn)
                nv.linear(xn, ws[i], out, tile_hint=32)
            row += f" LN+bn32 pair: {graph_timeit(pair, n_rot):6.2f}us"
        print(row, flush=True)


def skinny2():
    """Dissect the decode GEMM: time vs M (A traffic), K (main-loop length), weights from L2 (rot=1) vs HBM."""
    print("decode GEMM dissection, in-graph us per launch: N=1280, bn32")
    for m in ((128,) if os.environ.get('SK2_FAST') else (16, 64, 128)):
        for k in (320, 640, 1280, 2560, 5120):
            row = f"M={m:4d} K={k:5d}: "
            for n_rot in (1, max(2, int(300e6 // (1280 * k * 2)) + 1)):
                ws = [torch.randn(1280, k, device="cuda").bfloat16() * 0.02 for _ in range(n_rot)]
                a = torch.randn(m, k, device="cuda").bfloat16()
                out = torch.empty(m, 1280, device="cuda", dtype=torch.bfloat16)
                for hint in (32, 64):
                    us = graph_timeit(lambda i: nv.linear(a, ws[i], out, tile_hint=hint), n_rot)
                    row += f" rot{n_rot:3d} bn{hint}: {us:6.2f}us |"
                del ws
            print(row, flush=True)


def skinny3():
    """(tile, cluster size) grid for the decode GEMM shapes; cluster size forced through WF_SKINNY_CS."""
    cs = os.environ.get("WF_SKINNY_CS", "auto")
    for (n, k) in ((1280, 1280), (5120, 1280), (1280, 5120), (3840, 1280)):
        n_rot = max(2, int(300e6 // (n * k * 2)) + 1)
        ws = [torch.randn(n, k, device="cuda").bfloat16() * 0.02 for _ in range(n_rot)]
        a = torch.randn(128, k, device="cuda").bfloat16()
        out = torch.empty(128, n, device="cuda", dtype=torch.bfloat16)
        row = f"cs={cs} N={n:5d} K={k:5d}: "
        for hint in (32, 64, 128, 256):
            try:
                us = graph_timeit(lambda i: nv.linear(a, ws[i], out, tile_hint=hint), n_rot)
                row += f" bn{hint}: {us:6.2f}us |"
            except Exception as e:  # noqa: BLE001
                row += f" bn{hint}:   n/a    |"
        print(row, flush=True)
        del ws


def gemm():
    print("skinny / decode GEMMs (bf16, M = batch rows): us, GB/s of weight bytes")
    for m in (16, 128):
        for (n, k) in ((1280, 1280), (2560, 1280), (5120, 1280), (1280, 5120), (51865, 1280), (768, 768), (3072, 768)):
            n_rot = max(2, int(300e6 // (n * k * 2)) + 1)
            ws = [torch.randn(n, k, device="cuda").bfloat16() * 0.02 for _ in range(n_rot)]
            a = torch.randn(m, k, device="cuda").bfloat16()
            out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
            row = f"M={m:4d} N={n:6d} K={k:5d}: "
            for hint in (32, 64, 128, 256):
                try:
                    us = timeit(lambda i: nv.linear(a, ws[i], out, tile_hint=hint), n_rot)
                    row += f" bn{hint}: {us:7.1f}us {n * k * 2 / us / 1e3:6.0f}GB/s |"
                except Exception as e:
                    row += f" bn{hint}: n/a |"
            print(row, flush=True)
    print("large GEMMs (encoder): us, TFLOP/s")
    for (m, n, k, kw) in ((192000, 3840, 1280, {}), (192000, 1280, 1280, {}), (192000, 5120, 1280, dict(act=1)),
                          (192000, 1280, 5120, {}), (24000, 2304, 768, {}), (24000, 3072, 768, dict(act=1))):
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = torch.randn(n, k, device="cuda").bfloat16() * 0.02
        bias = torch.randn(n, device="cuda")
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        row = f"M={m} N={n} K={k} {kw}: "
        for hint in (128, 256):
            us = timeit(lambda i: nv.linear(a, w, out, bias=bias, tile_hint=hint, **kw), 1, iters=5)
            row += f" bn{hint}: {us:8.1f}us {2 * m * n * k / us / 1e6:7.1f}TF |"
        print(row, flush=True)


def gemm2():
    """Epilogue sensitivity of the encoder MLP-up GEMM (M=192000, N=5120, K=1280)."""
    m, n, k = 192000, 5120, 1280
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = torch.randn(n, k, device="cuda").bfloat16() * 0.02
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    out32 = torch.empty(m, n // 4, device="cuda", dtype=torch.float32)
    for name, fn in (("no bias, no act", lambda i: nv.linear(a, w, out)),
                     ("bias", lambda i: nv.linear(a, w, out, bias=bias)),
                     ("bias + gelu", lambda i: nv.linear(a, w, out, bias=bias, act=1)),
                     ("N=1280 slice, fp32 out", lambda i: nv.linear(a, w[:1280], out32))):
        us = timeit(fn, 1, iters=5)
        nn = 1280 if "slice" in name else n
        print(f"{name:28s}: {us:8.1f}us {2 * m * nn * k / us / 1e6:7.1f}TF", flush=True)


def attn():
    print("decode attention (bf16): us, GB/s of K+V bytes")
    for (B, H, Tk, G) in ((128, 20, 1500, 1), (128, 20, 750, 1), (16, 12, 1500, 1), (64, 16, 1500, 5), (1, 6, 1500, 1)):
        d = H * 64
        n_rot = max(2, int(400e6 // (B * Tk * 2 * d * 2)) + 1)
        kvs = [torch.randn(B * Tk, 2 * d, device="cuda").bfloat16() for _ in range(n_rot)]
        q = torch.randn(B * G, d, device="cuda").bfloat16()
        out = torch.empty_like(q)
        ws = torch.empty(nv.attention_decode_workspace_bytes(B * G, H), dtype=torch.uint8, device="cuda")
        row = f"B={B} H={H} Tk={Tk} G={G}: "
        for splits in (0, 1, 2, 4, 6, 12):
            os.environ["WF_DECODE_SPLITS"] = str(splits)
            us = timeit(lambda i: nv.attention_decode(q, kvs[i][:, :d], kvs[i][:, d:], 2 * d, Tk * 2 * d, 64, out, G, H,
                                                      None, 0, Tk, ws), n_rot)
            hm = [kv.view(B, Tk, 2 * H, 64) for kv in kvs]  # same bytes viewed head-major [B, 2H, T, 64]
            us_hm = timeit(lambda i: nv.attention_decode(q, hm[i].view(B, 2 * H, Tk, 64), hm[i].view(B, 2 * H, Tk, 64)[:, H:], 64,
                                                         2 * H * Tk * 64, Tk * 64, out, G, H, None, 0, Tk, ws), n_rot)
            row += f" [hm {us_hm:7.1f}us {B * Tk * 2 * d * 2 / us_hm / 1e3:6.0f}GB/s]"
            row += f" s{splits}: {us:7.1f}us {B * Tk * 2 * d * 2 / us / 1e3:6.0f}GB/s |"
        os.environ["WF_DECODE_SPLITS"] = "0"
        print(row, flush=True)


def attn2():
    """Head-major decode attention, persistent kernel vs the per-item kernel (WF_DECODE_PERSIST), in-graph."""
    persist = os.environ.get("WF_DECODE_PERSIST", "1")
    for (B, H, Tk, G) in ((128, 20, 1500, 1), (128, 20, 750, 1), (64, 20, 1500, 1), (64, 20, 750, 1), (64, 16, 1500, 5)):
        d = H * 64
        n_rot = max(2, int(600e6 // (B * Tk * 2 * d * 2)) + 1)
        kvs = [torch.randn(B, 2 * H, Tk, 64, device="cuda").bfloat16() for _ in range(n_rot)]
        q = torch.randn(B * G, d, device="cuda").bfloat16()
        out = torch.empty_like(q)
        ws = torch.empty(nv.attention_decode_workspace_bytes(B * G, H), dtype=torch.uint8, device="cuda")
        us = graph_timeit(lambda i: nv.attention_decode(q, kvs[i], kvs[i][:, H:], 64, 2 * H * Tk * 64, Tk * 64, out, G, H,
                                                        None, 0, Tk, ws), n_rot, reps=16)
        print(f"persist={persist} B={B} H={H} Tk={Tk} G={G}: {us:7.1f}us {B * Tk * 2 * d * 2 / us / 1e3:6.0f}GB/s", flush=True)
        del kvs


def ln():
    for rows, d in ((128, 1280), (192000, 1280), (24000, 768)):
        x = torch.randn(rows, d, device="cuda").bfloat16()
        w, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
        out = torch.empty_like(x)
        us = timeit(lambda i: nv.layernorm(x, w, b, out), 1)
        print(f"layernorm rows={rows} d={d}: {us:.1f} us, {rows * d * 4 / us / 1e3:.0f} GB/s")


def mel():
    import whisper
    for B in (1, 16, 128, 1024):
        pcm = torch.randn(B, 480000, device="cuda") * 0.1
        for n_mels in (80, 128):
            us = timeit(lambda i: whisper.log_mel_spectrogram(pcm, n_mels=n_mels, per_clip_max=True), 1, iters=5)
            by = B * (480000 * 4 + n_mels * 3000 * 4)
            print(f"logmel B={B} n_mels={n_mels}: {us:.1f} us, {by / us / 1e3:.0f} GB/s, {B / us * 1e6:.0f} clips/s")


def fa():
    for (B, H, T) in ((128, 20, 1500), (16, 12, 1500)):
        d = H * 64
        qkv = torch.randn(B * T, 3 * d, device="cuda").bfloat16()
        out = torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda i: nv.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, T, T, H, False), 1, iters=5)
        print(f"attention_full B={B} H={H} T={T}: {us:.1f} us, {4 * B * H * T * T * 64 / us / 1e6:.1f} TFLOP/s")


def latent():
    """Absorbed cross-attention (q' -> stream the source rows once -> value projection) vs the K/V-cache kernel."""
    print("latent cross-attention, in-graph us per launch; GB/s of source bytes (B*T*d*2)")
    shapes = ((128, 20, 1500), (128, 20, 750), (64, 16, 1500), (16, 12, 1500))
    for (B, H, T) in shapes[:int(os.environ.get("LATENT_SHAPES", len(shapes)))]:
        d = H * 64
        n_rot = 2 if B * T * d * 2 > 200e6 else 4
        srcs = [torch.randn(B, T, d, device="cuda").bfloat16() for _ in range(n_rot)]
        q = torch.randn(B, d, device="cuda").bfloat16()
        wk_t = (torch.randn(d, d, device="cuda") * (2.0 / d ** 0.5)).bfloat16()
        wv = (torch.randn(d, d, device="cuda") / d ** 0.5).bfloat16()
        bv = torch.zeros(d, device="cuda")
        qp = torch.empty(B, H, d, device="cuda", dtype=torch.bfloat16)
        ctx = torch.empty(B, H, d, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(B, d, device="cuda", dtype=torch.bfloat16)
        nv.latent_query(q, wk_t, qp, H)
        t_q = graph_timeit(lambda i: nv.latent_query(q, wk_t, qp, H), 1)
        t_a = graph_timeit(lambda i: nv.latent_attention(qp, srcs[i], ctx, H), n_rot, reps=16)
        t_s = None
        if nv.latent_split_supported(H):
            ctx2 = torch.zeros(2, B, H, d, device="cuda", dtype=torch.bfloat16)
            ml = torch.zeros(2, B, 32, 2, device="cuda")
            nv.latent_attention(qp, srcs[0], ctx2, H, ml=ml)
            t_s = graph_timeit(lambda i: nv.latent_attention(qp, srcs[i], ctx2, H, ml=ml), n_rot, reps=16)
            t_v2 = graph_timeit(lambda i: nv.latent_value(ctx2, wv, bv, out, H, ml=ml), 1)
        t_v = graph_timeit(lambda i: nv.latent_value(ctx, wv, bv, out, H), 1)
        split = "" if t_s is None else (f" | split form: attention {t_s:.1f} us = {B * T * d * 2 / t_s / 1e3:.0f} GB/s, "
                                        f"value {t_v2:.2f} us")
        print(f"B={B} H={H} T={T}: query {t_q:.2f} us | attention {t_a:.1f} us = {B * T * d * 2 / t_a / 1e3:.0f} GB/s "
              f"| value {t_v:.2f} us{split}", flush=True)


if __name__ == "__main__":
    for what in sys.argv[1:] or ["gemm", "attn", "ln", "mel", "fa"]:  # also: skinny, skinny2
        {"gemm": gemm, "attn": attn, "ln": ln, "mel": mel, "fa": fa, "skinny": skinny, "skinny2": skinny2, "skinny3": skinny3, "attn2": attn2, "gemm2": gemm2, "latent": latent}[what]()
