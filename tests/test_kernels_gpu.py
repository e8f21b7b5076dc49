"""GPU parity tests, kernel level: every libwf entry point against a plain PyTorch fp32 (or numpy fp64)
restatement of the same op, through the C ABI (ctypes)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, SAMPLE_IDX
from oracle import mel as omel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nv():
    from whisper import _native
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    _native.load()
    return _native


def _randn(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", dtype)


# ----------------------------------------------------------------------------- log-mel (tolerance: 1e-4 abs, north star)
@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_gaussian_and_chirp_vs_oracle_and_golden(nv, n_mels):
    import whisper
    from whisper._synthetic import synthetic_pcm
    g = np.load(f"{GOLDEN}/mel.npz")
    pcm = synthetic_pcm(2, seed=1234)
    got = whisper.log_mel_spectrogram(pcm[0].cuda(), n_mels=n_mels)
    assert got.shape == (n_mels, 3000) and got.dtype == torch.float32 and got.is_cuda
    want = omel.log_mel_spectrogram(pcm[0].numpy(), n_mels)
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-4
    idx = SAMPLE_IDX[SAMPLE_IDX < want.size]
    assert np.abs(got.cpu().numpy().reshape(-1)[idx] - g[f"gauss{n_mels}_samples"]).max() <= 1e-4
    chirp = torch.from_numpy(omel.chirp_kat()).cuda()
    c = whisper.log_mel_spectrogram(chirp, n_mels=n_mels).cpu().numpy()
    # 80 dB dynamic range: two correct fp32 implementations differ by ~1e-4 in the clamped floor bins, so the
    # gate is against the fp64 oracle with the reference's own error (6.3e-5) + 5e-5 margin
    assert np.abs(c - omel.log_mel_spectrogram(omel.chirp_kat(), n_mels)).max() <= 1.2e-4
    assert abs((c.max() - c.min()) - 2.0) < 1e-5  # max-8 clamp binds: min == max - 8/4
    assert np.abs(c.reshape(-1)[idx] - g[f"chirp{n_mels}_samples"]).max() <= 2e-4


_VARIANT_SCRIPT = r"""
import os, sys, math
import numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import whisper
from whisper import _native as nv
from oracle import mel as omel
what = sys.argv[3]
if what == "logmel":
    g = torch.Generator().manual_seed(5)
    pcm = torch.randn(3, 480000, generator=g) * 0.1
    worst = 0.0
    for n_mels in (80, 128):
        got = whisper.log_mel_spectrogram(pcm.cuda(), n_mels=n_mels, per_clip_max=True).cpu().numpy()
        want = omel.log_mel_spectrogram(pcm.numpy(), n_mels, per_clip_max=True)
        worst = max(worst, float(np.abs(got - want).max()))
        c = whisper.log_mel_spectrogram(torch.from_numpy(omel.chirp_kat()).cuda(), n_mels=n_mels).cpu().numpy()
        worst = max(worst, float(np.abs(c - omel.log_mel_spectrogram(omel.chirp_kat(), n_mels)).max()) / 1.2)
    print("RESULT", worst)
else:
    H, B, T = 20, 90, 300
    d = 64 * H
    g = torch.Generator().manual_seed(9)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).cuda()
    bf = torch.bfloat16
    q, src = r(B, d).to(bf), r(B, T, d).to(bf)
    wk, wv, bv = r(d, d, sc=2.0 / math.sqrt(d)).to(bf), r(d, d, sc=1.0 / math.sqrt(d)).to(bf), r(d, sc=0.1)
    qp = torch.empty(B, H, d, dtype=bf, device="cuda")
    ctx = torch.zeros(2, B, H, d, dtype=bf, device="cuda")
    ml = torch.zeros(2, B, 32, 2, device="cuda")
    out = torch.empty(B, d, dtype=bf, device="cuda")
    nv.latent_query(q, wk.t().contiguous(), qp, H)
    nv.latent_attention(qp, src, ctx, H, ml=ml)
    nv.latent_value(ctx, wv, bv, out, H, ml=ml)
    k = (src.float() @ wk.float().T).view(B, T, H, 64).permute(0, 2, 1, 3)
    v = (src.float() @ wv.float().T + bv).view(B, T, H, 64).permute(0, 2, 1, 3)
    w = torch.softmax((q.float().view(B, H, 1, 64) @ k.transpose(-1, -2)) * 0.125, dim=-1)
    ref = (w @ v).reshape(B, d)
    print("RESULT", float((out.float() - ref).norm() / ref.norm()))
"""


@pytest.mark.parametrize("env,what,bound", [("WF_LOGMEL_V1", "logmel", 1e-4), ("WF_LOGMEL_V2", "logmel", 1e-4),
                                            ("WF_LATENT_VALUE_TMA=0", "latent", 1.5e-2)])
def test_selectable_kernel_variants_stay_correct(nv, env, what, bound):
    """The A/B variants kept in the library (round-1 log-mel kernel, the 20-warp form of the new layout, the cp.async
    value projection) are chosen once per process through the environment: run each in a subprocess against the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    key, _, val = env.partition("=")
    e[key] = val or "1"
    r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, os.path.join(root, "whisper-flamingo_b200"), root, what],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    err = float([l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()[1])
    assert err <= bound, (env, err)


def test_logmel_batched_semantics_padding_and_cpu_input(nv):
    import whisper
    from whisper._synthetic import synthetic_pcm
    g = np.load(f"{GOLDEN}/mel.npz")
    pcm = synthetic_pcm(2, seed=1234)
    two = torch.stack([pcm[0], pcm[1] * 1e-3])
    b = whisper.log_mel_spectrogram(two.cuda(), n_mels=80).cpu().numpy()       # reference: global max
    assert np.abs(b - omel.log_mel_spectrogram(two.numpy(), 80)).max() <= 1e-4
    half = b.size // 2
    idx = SAMPLE_IDX[SAMPLE_IDX < half]
    got = np.concatenate([b.reshape(-1)[idx], b.reshape(-1)[half + idx]])
    assert np.abs(got - g["batch2_global_samples"]).max() <= 1e-4
    pc = whisper.log_mel_spectrogram(two.cuda(), n_mels=80, per_clip_max=True).cpu().numpy()
    assert np.abs(pc - omel.log_mel_spectrogram(two.numpy(), 80, per_clip_max=True)).max() <= 1e-4
    assert np.abs(pc[1] - b[1]).max() > 0.1
    # padding + short ragged clip + numpy / CPU input staged through the GPU, result returned on CPU
    sp = whisper.log_mel_spectrogram(pcm[0][:16000].numpy(), n_mels=80, padding=4800)
    assert not sp.is_cuda and sp.shape == (80, 130)
    assert np.abs(sp.numpy() - g["short_pad_full"]).max() <= 1e-4
    odd = whisper.log_mel_spectrogram(pcm[0][:12345].cuda(), n_mels=80).cpu().numpy()   # 77 frames, ragged tail
    assert odd.shape == (80, 77)
    assert np.abs(odd - omel.log_mel_spectrogram(pcm[0][:12345].numpy(), 80)).max() <= 1e-4
    assert whisper.log_mel_spectrogram(torch.zeros(80, 7)).shape == (80, 7)  # reference passthrough quirk
    with pytest.raises(AssertionError):
        whisper.log_mel_spectrogram(pcm[0].cuda(), n_mels=64)
    with pytest.raises(RuntimeError):
        whisper.log_mel_spectrogram(torch.zeros(100).cuda())  # reflect pad needs > 200 samples


def test_logmel_large_batch_linearity_property(nv):
    """Size-independent property at bench scale: scaling the PCM by 10 shifts log-mel by exactly 2/4 = 0.5."""
    import whisper
    from whisper._synthetic import synthetic_pcm
    pcm = synthetic_pcm(64, seed=7).cuda()
    a = whisper.log_mel_spectrogram(pcm, n_mels=80, per_clip_max=True)
    b = whisper.log_mel_spectrogram(pcm * 10.0, n_mels=80, per_clip_max=True)
    assert (b - a - 0.5).abs().max().item() <= 2e-5
    one = whisper.log_mel_spectrogram(pcm[17], n_mels=80)
    assert torch.equal(one, a[17])  # batched per-clip result is bit-identical to the single-clip call


# ----------------------------------------------------------------------------- linear layers
def _ref_linear(a, w, bias=None, act=0, gate=None, residual=None, res_row_mod=0):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if act == 1:
        v = F.gelu(v)
    if gate is not None:
        v = v * torch.tanh(gate)
    if residual is not None:
        r = residual.float()
        if res_row_mod:
            r = r[torch.arange(v.shape[0], device=v.device) % res_row_mod]
        v = v + r
    return v


@pytest.mark.parametrize("m,n,k,hint", [
    (128, 256, 64, 0), (300, 384, 240, 0), (3000, 1280, 1280, 256), (1500, 1152, 768, 128), (16, 1280, 1280, 0),
    (5, 51865, 384, 0), (129, 72, 200, 64), (640, 5120, 1280, 256), (100, 1280, 5120, 32),
])
def test_linear_bf16_tcgen05_vs_torch(nv, m, n, k, hint):
    a = _randn(m, k, dtype=torch.bfloat16, seed=1)
    w = _randn(n, k, dtype=torch.bfloat16, seed=2, scale=0.05)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    nv.linear(a, w, out, tile_hint=hint)
    want = _ref_linear(a, w)
    torch.cuda.synchronize()
    err = (out.float() - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), err  # one bf16 rounding of the output
    # fp32 output (logits path) is exact up to accumulation order
    out32 = torch.empty((m, n), dtype=torch.float32, device="cuda")
    nv.linear(a, w, out32, tile_hint=hint)
    assert (out32 - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


def test_linear_bf16_epilogues_and_strided_views(nv):
    m, n, k = 200, 384, 384
    a = _randn(m, k, dtype=torch.bfloat16, seed=3)
    w = _randn(n, k, dtype=torch.bfloat16, seed=4, scale=0.05)
    bias = _randn(n, seed=5)
    gate = torch.tensor([0.5], device="cuda")
    res = _randn(m, n, dtype=torch.bfloat16, seed=6)
    pos = _randn(50, n, dtype=torch.bfloat16, seed=7)
    for kw, ref_kw in [
        (dict(bias=bias), dict(bias=bias)),
        (dict(bias=bias, act=nv.ACT_GELU), dict(bias=bias, act=1)),
        (dict(bias=bias, residual=res), dict(bias=bias, residual=res)),
        (dict(bias=bias, act=nv.ACT_GELU, gate=gate, residual=res), dict(bias=bias, act=1, gate=gate, residual=res)),
        (dict(bias=bias, act=nv.ACT_GELU, residual=pos, res_row_mod=50), dict(bias=bias, act=1, residual=pos, res_row_mod=50)),
    ]:
        out = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
        nv.linear(a, w, out, **kw)
        want = _ref_linear(a, w, **ref_kw)
        assert (out.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item()), kw.keys()
    # in-place residual (x += f(x_n)) and column-sliced operands / outputs
    x = res.clone()
    nv.linear(a, w, x, bias=bias, residual=x)
    assert (x.float() - _ref_linear(a, w, bias=bias, residual=res)).abs().max().item() <= 3e-2 * 4
    big = _randn(m, 3 * k, dtype=torch.bfloat16, seed=8)
    outbig = torch.zeros((m, 2 * n), dtype=torch.bfloat16, device="cuda")
    nv.linear(big[:, k:2 * k], w, outbig[:, n:])
    assert (outbig[:, n:].float() - _ref_linear(big[:, k:2 * k], w)).abs().max().item() <= 3e-2 * 4
    assert outbig[:, :n].abs().max().item() == 0
    # KV-cache append: output rows land at column offset p * c_off_mul read from device memory
    cache = torch.zeros((m, 7, n), dtype=torch.bfloat16, device="cuda")
    p = torch.tensor([3], dtype=torch.int32, device="cuda")
    nv.linear(a, w, cache.view(m, 7 * n)[:, :n], bias=bias, c_off_ptr=p, c_off_mul=n)
    assert (cache[:, 3].float() - _ref_linear(a, w, bias=bias)).abs().max().item() <= 3e-2 * 4
    assert cache[:, :3].abs().max().item() == 0 and cache[:, 4:].abs().max().item() == 0


@pytest.mark.parametrize("m,n,k", [(128, 1280, 1280), (128, 1280, 5120), (16, 768, 3072), (130, 384, 1536), (5, 2560, 1280)])
def test_linear_bf16_split_k_vs_torch(nv, m, n, k):
    """Decode-shape GEMMs with the split-K workspace: partial tiles reduced by the last-arriving CTA; the
    arrival counters must be back to zero afterwards (the same workspace is reused by every launch)."""
    a = _randn(m, k, dtype=torch.bfloat16, seed=1)
    w = _randn(n, k, dtype=torch.bfloat16, seed=2, scale=0.05)
    bias, res = _randn(n, seed=3), _randn(m, n, dtype=torch.bfloat16, seed=4)
    gate = torch.tensor([0.5], device="cuda")
    ws = torch.zeros(4096 + 8 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    want = _ref_linear(a, w, bias=bias, act=1, gate=gate, residual=res)
    for _ in range(3):
        out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
        nv.linear(a, w, out, bias=bias, act=nv.ACT_GELU, gate=gate, residual=res, ws=ws)
        assert (out.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
        assert ws[:4096].view(torch.int32).abs().max().item() == 0
    out32 = torch.empty((m, n), dtype=torch.float32, device="cuda")
    nv.linear(a, w, out32, ws=ws)
    ref = _ref_linear(a, w)
    assert (out32 - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("m,n,k,hint", [
    (128, 1280, 1280, 0), (128, 1280, 1280, 32), (128, 1280, 1280, 64), (128, 1280, 1280, 128), (128, 1280, 1280, 256),
    (128, 5120, 1280, 0), (128, 1280, 5120, 0), (128, 3840, 1280, 0), (64, 768, 768, 0), (16, 3072, 768, 0),
    (1, 384, 384, 0), (100, 1536, 384, 0), (37, 1000, 200, 0), (128, 4096, 1024, 0), (128, 1024, 4096, 64),
])
def test_linear_bf16_skinny_cluster_split_k_vs_torch(nv, m, n, k, hint):
    """Decode-step GEMMs (M <= 128): thread-block cluster split-K with the DSMEM reduction, every epilogue option,
    odd sizes (K and N tails are TMA zero-fill / predicated stores)."""
    a = _randn(m, k, dtype=torch.bfloat16, seed=1)
    w = _randn(n, k, dtype=torch.bfloat16, seed=2, scale=0.05)
    bias, res = _randn(n, seed=3), _randn(m, n, dtype=torch.bfloat16, seed=4)
    gate = torch.tensor([0.5], device="cuda")
    want = _ref_linear(a, w, bias=bias, act=1, gate=gate, residual=res)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    nv.linear(a, w, out, bias=bias, act=nv.ACT_GELU, gate=gate, residual=res, tile_hint=hint)
    assert (out.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
    out32 = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
    nv.linear(a, w, out32, tile_hint=hint)
    ref = _ref_linear(a, w)
    assert (out32 - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    # in-place residual, repeated launches (no state carried between launches)
    x = res.clone()
    for _ in range(2):
        nv.linear(a, w, x, bias=bias, residual=x, tile_hint=hint)
    want2 = _ref_linear(a, w, bias=bias, residual=_ref_linear(a, w, bias=bias, residual=res).bfloat16())
    assert (x.float() - want2).abs().max().item() <= 5e-2 * max(1.0, want2.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(128, 1280, 1280), (128, 5120, 1280), (16, 768, 768), (3, 384, 384), (128, 3840, 1280)])
def test_linear_bf16_fused_layernorm_vs_torch(nv, m, n, k):
    """LayerNorm (reference model.py:30-32, fp32 statistics) folded into the following Linear: the kernel sees the raw
    rows, gamma is folded into W, beta into the bias, and mean / rstd come from the staged operand tiles."""
    x = (_randn(m, k, seed=1) * 1.7 + 0.6).bfloat16()           # non-zero mean: the mean correction must matter
    gamma, beta = 1.0 + 0.2 * _randn(k, seed=2), 0.3 * _randn(k, seed=3)
    w32 = _randn(n, k, seed=4, scale=0.05)
    bias = _randn(n, seed=5)
    w_fold = (w32 * gamma[None, :]).bfloat16()
    colsum = w_fold.float().sum(1).contiguous()
    bias_fold = (bias + w32 @ beta).contiguous()
    want = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w32, bias)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    nv.linear(x, w_fold, out, bias=bias_fold, ln_colsum=colsum, ln_eps=1e-5)
    scale = max(1.0, want.abs().max().item())
    assert (out.float() - want).abs().max().item() <= 3e-2 * scale
    # against the unfused sequence of this library (LayerNorm kernel -> bf16 -> GEMM): same quantity, other rounding
    xn = torch.empty_like(x)
    nv.layernorm(x, gamma, beta, xn)
    out_unfused = torch.empty_like(out)
    nv.linear(xn, w32.bfloat16(), out_unfused, bias=bias)
    assert (out.float() - out_unfused.float()).abs().max().item() <= 4e-2 * scale
    # with GELU + fp32 output
    out32 = torch.empty((m, n), dtype=torch.float32, device="cuda")
    nv.linear(x, w_fold, out32, bias=bias_fold, act=nv.ACT_GELU, ln_colsum=colsum, ln_eps=1e-5)
    assert (out32 - F.gelu(want)).abs().max().item() <= 2e-2 * scale


@pytest.mark.parametrize("m,d,n", [(300, 384, 1152), (1500, 1280, 5120), (129, 768, 768), (4100, 1280, 3840)])
def test_linear_bf16_layernorm_statistics_across_gemms(nv, m, d, n):
    """Encoder-sized problems (M > 128; M >= 2048 runs on CTA pairs): the GEMM that writes the residual stream emits per-row (sum, sum of squares)
    partials, the next GEMM applies the folded LayerNorm from them (reference model.py:30-32 + :35-41)."""
    a = _randn(m, d, dtype=torch.bfloat16, seed=1)
    w0 = _randn(d, d, dtype=torch.bfloat16, seed=2, scale=0.05)
    res = (_randn(m, d, seed=3) + 0.5).bfloat16()
    slots = 2 * ((d + 255) // 256)
    stats = torch.full((m, slots, 2), float("nan"), dtype=torch.float32, device="cuda")
    x = res.clone()
    nv.linear(a, w0, x, residual=x, tile_hint=256, stat_out=stats)       # x = res + a w0^T, statistics of the stored x
    xf = x.float()
    assert (stats[:, :, 0].sum(1) - xf.sum(1)).abs().max().item() <= 2e-2
    assert (stats[:, :, 1].sum(1) - (xf * xf).sum(1)).abs().max().item() <= 1e-3 * (xf * xf).sum(1).max().item()
    gamma, beta = 1.0 + 0.2 * _randn(d, seed=4), 0.3 * _randn(d, seed=5)
    w32, bias = _randn(n, d, seed=6, scale=0.05), _randn(n, seed=7)
    w_fold = (w32 * gamma[None, :]).bfloat16()
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    nv.linear(x, w_fold, out, bias=(bias + (w32 * beta[None, :]).sum(1)).contiguous(), act=nv.ACT_GELU,
              ln_colsum=w_fold.float().sum(1).contiguous(), ln_eps=1e-5, stat_in=stats)
    want = F.gelu(F.linear(F.layer_norm(xf, (d,), gamma, beta, 1e-5), w32, bias))
    assert (out.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
    if m <= 512:
        # no statistics given: up to 512 rows (beam-search steps) the row-tiled decode GEMM computes the LayerNorm itself
        out2 = torch.full_like(out, float("nan"))
        nv.linear(x, w_fold, out2, bias=(bias + (w32 * beta[None, :]).sum(1)).contiguous(), act=nv.ACT_GELU,
                  ln_colsum=w_fold.float().sum(1).contiguous(), ln_eps=1e-5)
        assert (out2.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
    else:
        with pytest.raises(nv.WfError):  # more than 512 rows and no statistics: refused, not silently wrong
            nv.linear(x, w_fold, out, ln_colsum=w_fold.float().sum(1).contiguous())


def test_linear_bf16_fused_qkv_two_destinations(nv):
    """q | k,v projection in one GEMM: q columns row-major, k/v columns appended to the head-major self-attention cache
    at the device-side position (reference model.py:76-85 computes the three projections separately)."""
    R, H, k, cap = 7, 6, 384, 20
    d = H * 64
    x = _randn(R, k, dtype=torch.bfloat16, seed=1)
    w = _randn(3 * d, k, dtype=torch.bfloat16, seed=2, scale=0.05)
    bias = _randn(3 * d, seed=3)
    q = torch.full((R, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    cache = torch.zeros((R, 2 * H, cap, 64), dtype=torch.bfloat16, device="cuda")
    pos = torch.tensor([11], dtype=torch.int32, device="cuda")
    nv.linear(x, w, q, bias=bias, out2=cache.view(-1, 64), split_n=d, head_major=(2 * H, cap, 1), c_off_ptr=pos,
              c_off_mul=64)
    want = _ref_linear(x, w, bias=bias)
    tol = 3e-2 * max(1.0, want.abs().max().item())
    assert (q.float() - want[:, :d]).abs().max().item() <= tol
    assert (cache[:, :, 11].float() - want[:, d:].view(R, 2 * H, 64)).abs().max().item() <= tol
    cache[:, :, 11] = 0
    assert cache.abs().max().item() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_head_major_kv_output(nv, dtype):
    """K/V projection written straight into a head-major cache [B, 2H, T, 64] (+ device-side append offset)."""
    B, T, H, k = 3, 50, 6, 384
    n = 2 * H * 64
    a = _randn(B * T, k, dtype=dtype, seed=1)
    w = _randn(n, k, dtype=dtype, seed=2, scale=0.05)
    bias = _randn(n, seed=3)
    cache = torch.zeros((B, 2 * H, T, 64), dtype=dtype, device="cuda")
    nv.linear(a, w, cache.view(-1, 64), bias=bias, head_major=(2 * H, T, T))
    want = _ref_linear(a, w, bias=bias).view(B, T, 2 * H, 64).permute(0, 2, 1, 3)
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    assert (cache.float() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    # decode-step append: one row per batch entry, position read from device memory
    R, cap = 5, 20
    a1 = _randn(R, k, dtype=dtype, seed=4)
    self_kv = torch.zeros((R, 2 * H, cap, 64), dtype=dtype, device="cuda")
    pos = torch.tensor([7], dtype=torch.int32, device="cuda")
    ws = torch.zeros(4096 + 1024 * 1024, dtype=torch.uint8, device="cuda") if dtype == torch.bfloat16 else None
    nv.linear(a1, w, self_kv.view(-1, 64), bias=bias, head_major=(2 * H, cap, 1), c_off_ptr=pos, c_off_mul=64, ws=ws)
    want1 = _ref_linear(a1, w, bias=bias).view(R, 2 * H, 64)
    assert (self_kv[:, :, 7].float() - want1).abs().max().item() <= tol * max(1.0, want1.abs().max().item())
    self_kv[:, :, 7] = 0
    assert self_kv.abs().max().item() == 0


@pytest.mark.parametrize("m,n,k", [(1, 384, 384), (70, 130, 50), (3000, 384, 240), (64, 51865, 384)])
def test_linear_f32_vs_torch(nv, m, n, k):
    a, w = _randn(m, k, seed=1), _randn(n, k, seed=2, scale=0.05)
    bias, gate, res = _randn(n, seed=3), torch.tensor([0.5], device="cuda"), _randn(m, n, seed=4)
    out = torch.empty((m, n), device="cuda")
    nv.linear(a, w, out, bias=bias, act=nv.ACT_GELU, gate=gate, residual=res)
    want = _ref_linear(a, w, bias=bias, act=1, gate=gate, residual=res)
    assert (out - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


# ----------------------------------------------------------------------------- row-wise kernels
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("d", [384, 768, 1024, 1280, 100])
def test_layernorm_vs_torch(nv, dtype, tol, d):
    x = _randn(77, d, dtype=dtype, seed=1, scale=3.0)
    w, b = 1 + 0.1 * _randn(d, seed=2), 0.1 * _randn(d, seed=3)
    out = torch.empty_like(x)
    nv.layernorm(x, w, b, out)
    want = F.layer_norm(x.float(), (d,), w, b, 1e-5)
    assert (out.float() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())


def test_im2col_embed_addrow_cast(nv):
    B, C, T = 2, 80, 301
    mel = _randn(B, C, T, seed=1)
    w = _randn(64, C, 3, seed=2, scale=0.1)
    a1 = torch.empty((B * T, 3 * C), device="cuda")
    nv.im2col_k3(mel, C * T, T, 1, B, C, T, 1, a1)
    got = (a1 @ w.reshape(64, -1).t()).view(B, T, 64).permute(0, 2, 1)
    assert (got - F.conv1d(mel, w, padding=1)).abs().max().item() <= 1e-4
    h = _randn(B * T, 64, seed=3)                                   # NWC activations, stride-2 conv
    w2 = _randn(32, 64, 3, seed=4, scale=0.1)
    T2 = (T - 1) // 2 + 1
    a2 = torch.empty((B * T2, 3 * 64), device="cuda")
    nv.im2col_k3(h, T * 64, 1, 64, B, 64, T, 2, a2)
    got = (a2 @ w2.reshape(32, -1).t()).view(B, T2, 32)
    want = F.conv1d(h.view(B, T, 64).permute(0, 2, 1), w2, stride=2, padding=1).permute(0, 2, 1)
    assert (got - want).abs().max().item() <= 1e-4
    a1b = torch.empty((B * T, 3 * C), dtype=torch.bfloat16, device="cuda")
    nv.im2col_k3(mel, C * T, T, 1, B, C, T, 1, a1b)
    assert torch.equal(a1b, a1.bfloat16())
    # embedding
    tok_emb, pos_emb = _randn(1000, 384, seed=5), _randn(448, 384, seed=6)
    toks = torch.randint(0, 1000, (3, 10), dtype=torch.int32, device="cuda")
    out = torch.empty((3 * 10, 384), device="cuda")
    nv.embed(toks, 10, None, 0, tok_emb, pos_emb, out, n_pos=10)
    assert torch.equal(out.view(3, 10, 384), tok_emb[toks.long()] + pos_emb[:10])
    pos = torch.tensor([4], dtype=torch.int32, device="cuda")
    out1 = torch.empty((3, 384), dtype=torch.bfloat16, device="cuda")
    nv.embed(toks, 10, pos, 0, tok_emb, pos_emb, out1)
    assert torch.equal(out1, (tok_emb[toks[:, 4].long()] + pos_emb[4]).bfloat16())
    # positional add + casts
    x = _randn(20, 384, seed=7)
    o = torch.empty((20, 384), dtype=torch.bfloat16, device="cuda")
    nv.add_rowmod(x, pos_emb, o, 5)
    assert torch.equal(o, (x + pos_emb[torch.arange(20, device="cuda") % 5]).bfloat16())
    assert torch.equal(nv.cast(x, torch.empty_like(x, dtype=torch.bfloat16)), x.bfloat16())


# ----------------------------------------------------------------------------- attention
def _ref_attention(q, k, v, B, Tq, Tk, H, causal):
    d = H * 64
    qf = q.float().view(B, Tq, H, 64).permute(0, 2, 1, 3)
    kf = k.float().view(B, Tk, H, 64).permute(0, 2, 1, 3)
    vf = v.float().view(B, Tk, H, 64).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * 0.125
    if causal:
        s = s + torch.full((Tq, Tk), float("-inf"), device=q.device).triu_(1)
    return (s.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B * Tq, d)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,Tq,Tk,H,causal", [(2, 1500, 1500, 6, False), (3, 37, 37, 6, True), (2, 8, 1500, 12, False),
                                              (2, 70, 100, 6, False), (1, 130, 130, 20, True), (3, 300, 700, 2, False),
                                              (1, 256, 256, 1, False), (2, 500, 1500, 20, False)])
def test_attention_full_vs_torch(nv, dtype, tol, B, Tq, Tk, H, causal):
    d = H * 64
    qkv = _randn(B * Tq, 3 * d, dtype=dtype, seed=1)
    kv = _randn(B * Tk, 2 * d, dtype=dtype, seed=2)
    q = qkv[:, :d]
    k, v = (qkv[:, d:2 * d], qkv[:, 2 * d:]) if Tq == Tk else (kv[:, :d], kv[:, d:])
    out = torch.full((B * Tq, d), float("nan"), dtype=dtype, device="cuda")
    nv.attention(q, k, v, out, B, Tq, Tk, H, causal)
    want = _ref_attention(q, k, v, B, Tq, Tk, H, causal)
    assert (out.float() - want).abs().max().item() <= tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,G,H,Tk,dyn", [(2, 1, 6, 1500, False), (16, 1, 12, 750, False), (3, 5, 16, 1500, False),
                                          (4, 1, 6, 37, True), (1, 1, 6, 300, True), (2, 3, 6, 100, False),
                                          # >= 2 items per SM: the persistent kernel (head-major bf16 only)
                                          (32, 1, 12, 750, False), (20, 5, 16, 1500, False), (64, 1, 6, 300, False),
                                          (50, 2, 6, 257, False),
                                          # beams sharing a cache: tensor-core multi-query kernel (head-major bf16)
                                          (4, 8, 6, 700, False), (6, 4, 12, 1500, False), (3, 6, 20, 128, False),
                                          # any group size: 7 in one pass, 10 / 15 / 17 in chunks of <= 8 queries
                                          (3, 7, 6, 1500, False), (2, 10, 12, 750, False), (2, 15, 6, 300, False),
                                          (2, 17, 6, 100, False), (2, 9, 6, 40, True)])
@pytest.mark.parametrize("head_major", [False, True])
def test_attention_decode_vs_torch(nv, dtype, tol, B, G, H, Tk, dyn, head_major):
    d, R, cap = H * 64, B * G, 448
    q = _randn(R, d, dtype=dtype, seed=1)
    T_alloc = cap if dyn else Tk
    kv = _randn(B * T_alloc, 2 * d, dtype=dtype, seed=2)   # row-interleaved reference layout [B*T, K | V]
    out = torch.full((R, d), float("nan"), dtype=dtype, device="cuda")
    ws = torch.empty(nv.attention_decode_workspace_bytes(R, H), dtype=torch.uint8, device="cuda")
    lp = torch.tensor([Tk - 1], dtype=torch.int32, device="cuda") if dyn else None
    len_args = (lp, 1, cap) if dyn else (None, 0, Tk)
    if head_major:  # [B, 2H, T, 64]: K heads then V heads
        hm = kv.view(B, T_alloc, 2 * H, 64).permute(0, 2, 1, 3).contiguous()
        nv.attention_decode(q, hm, hm[:, H:], 64, 2 * H * T_alloc * 64, T_alloc * 64, out, G, H, *len_args, ws)
    else:
        nv.attention_decode(q, kv[:, :d], kv[:, d:], 2 * d, T_alloc * 2 * d, 64, out, G, H, *len_args, ws)
    kvv = kv.view(B, T_alloc, 2 * d)[:, :Tk]
    k = kvv[..., :d].repeat_interleave(G, 0).reshape(R * Tk, d)
    v = kvv[..., d:].repeat_interleave(G, 0).reshape(R * Tk, d)
    want = _ref_attention(q, k, v, R, 1, Tk, H, False)
    assert (out.float() - want).abs().max().item() <= tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("R,H,cap,length", [(10, 6, 32, 17), (130, 16, 70, 70), (320, 16, 69, 41), (7, 20, 448, 300)])
def test_attention_decode_paged_row_table_vs_torch(nv, dtype, tol, R, H, cap, length):
    """wf_attention_decode_paged (beam search: reference rearrange_kv_cache, decoding.py:173-180, done by rewriting a
    [rows, T_cap] table instead of moving K/V): key j of row r is read from cache entry row_table[r, j].  Checked with a
    random table per position against a torch gather, head-major cache [R, 2H, T_cap, 64] with the length in device
    memory - the layout and call the R > 128 beam session (BASELINE config 3) makes."""
    d = H * 64
    q = _randn(R, d, dtype=dtype, seed=11)
    cache = _randn(R, 2 * H, cap, 64, dtype=dtype, seed=12)
    g = torch.Generator(device="cpu").manual_seed(13)
    table = torch.randint(0, R, (R, cap), generator=g, dtype=torch.int32).cuda()
    out = torch.full((R, d), float("nan"), dtype=dtype, device="cuda")
    ws = torch.empty(nv.attention_decode_workspace_bytes(R, H), dtype=torch.uint8, device="cuda")
    lp = torch.tensor([length - 1], dtype=torch.int32, device="cuda")
    nv.attention_decode(q, cache, cache[:, H:], 64, 2 * H * cap * 64, cap * 64, out, 1, H, lp, 1, cap, ws,
                        row_table=table)
    # gathered[r, :, j] = cache[table[r, j], :, j]
    idx = table[:, :length].long()
    pos = torch.arange(length, device="cuda")[None, :].expand(R, -1)
    gathered = cache.float()[idx, :, pos]                       # [R, length, 2H, 64]
    k = gathered[:, :, :H].reshape(R * length, d)
    v = gathered[:, :, H:].reshape(R * length, d)
    want = _ref_attention(q, k, v, R, 1, length, H, False)
    assert (out.float() - want).abs().max().item() <= tol
    # identity table == the un-paged call
    ident = torch.arange(R, dtype=torch.int32, device="cuda")[:, None].repeat(1, cap).contiguous()
    out2 = torch.empty_like(out)
    nv.attention_decode(q, cache, cache[:, H:], 64, 2 * H * cap * 64, cap * 64, out2, 1, H, lp, 1, cap, ws,
                        row_table=ident)
    out3 = torch.empty_like(out)
    nv.attention_decode(q, cache, cache[:, H:], 64, 2 * H * cap * 64, cap * 64, out3, 1, H, lp, 1, cap, ws)
    assert (out2.float() - out3.float()).abs().max().item() <= tol


# ----------------------------------------------------------------------------- sampling
def test_sampler_uniform_is_strictly_inside_unit_interval(nv):
    """The Gumbel-max sampler draws u with a counter RNG; u == 1 (or 0) makes -log(-log u) infinite and that token wins
    whatever its logit (24 random bits + 0.5 rounds up to 2^24 in fp32).  2^28 draws hit the extreme
    buckets of a 23-bit uniform ~32 times each: min / max must be exactly the ends of [2^-24, 1 - 2^-24]."""
    lo, hi = nv.debug_uniform_range(1234, 1 << 28, "cuda")
    assert 0.0 < lo and hi < 1.0, (lo, hi)
    assert lo == 2.0 ** -24 and hi == 1.0 - 2.0 ** -24, (lo, hi)
    assert math.isfinite(-math.log(-math.log(hi))) and math.isfinite(-math.log(-math.log(lo)))


def test_sample_greedy_and_topk_vs_torch(nv):
    R, V, eot, nosp = 6, 51865, 50257, 50362
    logits = _randn(R, V + 7, seed=1, scale=2.0)
    sup = torch.zeros(V, dtype=torch.uint8, device="cuda")
    sup[[5, 17, nosp, 50258]] = 1
    first = torch.zeros(V, dtype=torch.uint8, device="cuda")
    first[[220, eot]] = 1
    logits[0, 5] = 100.0      # suppressed max must be ignored
    logits[1, 220] = 90.0     # blank is masked on the first sampled step only
    logits[2, eot] = 80.0     # EOT masked on the first step
    n_init = 4
    tokens = torch.zeros((R, 16), dtype=torch.int32, device="cuda")
    tokens[:, :n_init] = torch.tensor([50258, 50259, 50359, 50363], dtype=torch.int32)
    state = torch.tensor([0, n_init, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device="cuda")
    slp = torch.zeros(R, device="cuda")
    nsp = torch.full((R,), float("nan"), device="cuda")
    # t = 0: SOT position -> only no_speech_prob is produced
    nv.sample_greedy(logits, V, sup, first, tokens, state, slp, nsp, eot, nosp)
    nv.step_advance(state, R)
    want_nsp = logits[:, :V].softmax(-1)[:, nosp]
    assert (nsp - want_nsp).abs().max().item() <= 1e-7 + 1e-4 * want_nsp.max().item()
    assert slp.abs().max().item() == 0 and tokens[:, n_init:].abs().max().item() == 0
    for _ in range(2):
        nv.sample_greedy(logits, V, sup, first, tokens, state, slp, nsp, eot, nosp)
        nv.step_advance(state, R)
    assert state[0].item() == 3
    # t = 3: first sampled token
    nv.sample_greedy(logits, V, sup, first, tokens, state, slp, nsp, eot, nosp)
    nv.step_advance(state, R)
    masked = logits[:, :V].clone()
    masked[:, sup.bool()] = -np.inf
    masked[:, first.bool()] = -np.inf
    want_tok = masked.argmax(-1)
    want_lp = masked.log_softmax(-1).gather(1, want_tok[:, None])[:, 0]
    assert tokens[:, n_init].tolist() == want_tok.tolist()
    assert (slp - want_lp).abs().max().item() <= 1e-4
    # t = 4: blank/EOT allowed again; make row 3 finish and row 4 already finished
    logits[3, eot] = 95.0
    tokens[4, n_init] = eot
    before = slp.clone()
    nv.sample_greedy(logits, V, sup, first, tokens, state, slp, nsp, eot, nosp)
    nv.step_advance(state, R)
    m2 = logits[:, :V].clone()
    m2[:, sup.bool()] = -np.inf
    want2 = m2.argmax(-1)
    want2[4] = eot
    assert tokens[:, n_init + 1].tolist() == want2.tolist()
    assert tokens[3, n_init + 1].item() == eot and tokens[1, n_init + 1].item() == 220
    assert slp[4].item() == before[4].item()  # finished rows stop accumulating
    assert state[2].item() == 0               # not all rows at EOT
    # ties resolve to the lowest index (torch.argmax convention)
    logits2 = torch.zeros((R, V + 7), device="cuda")
    nv.sample_greedy(logits2, V, sup, None, tokens, state, slp, nsp, eot, nosp)
    assert tokens[0, n_init + 2].item() == 0 and tokens[4, n_init + 2].item() == eot
    # top-k log-probabilities (beam search)
    k = 6
    vals = torch.empty((R, k), device="cuda")
    idx = torch.empty((R, k), dtype=torch.int32, device="cuda")
    nv.topk_logprobs(logits, V, sup, first, None, n_init, n_init + 3, eot, (-1, -1, -1), k, vals, idx)
    tv, ti = m2.log_softmax(-1).topk(k)
    assert idx.tolist() == ti.tolist() and (vals - tv).abs().max().item() <= 1e-4


def test_sample_temperature_follows_softmax(nv):
    """temperature > 0 (reference decoding.py:286-291: Categorical(logits / T).sample(), sum_logprobs from the
    UN-tempered log_softmax): Gumbel-max with a counter RNG must reproduce softmax(filtered logits / T).  The random
    stream is not torch's, so the check is statistical: 8192 rows sharing one logits row, 5-sigma binomial bounds."""
    R, V, eot, nosp, T = 8192, 51865, 50257, 50362, 0.7
    row = _randn(1, V + 7, seed=3, scale=2.0)
    row[0, 1000:1006] += 9.0  # a handful of likely tokens so that the empirical frequencies are measurable
    logits = row.expand(R, -1).contiguous()
    sup = torch.zeros(V, dtype=torch.uint8, device="cuda")
    sup[[1002, nosp]] = 1
    n_init = 2
    tokens = torch.zeros((R, 8), dtype=torch.int32, device="cuda")
    tokens[:, :n_init] = torch.tensor([50258, 50363], dtype=torch.int32)
    state = torch.tensor([n_init - 1, n_init, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device="cuda")
    slp = torch.zeros(R, device="cuda")
    nsp = torch.zeros(R, device="cuda")
    nv.sample_greedy(logits, V, sup, None, tokens, state, slp, nsp, eot, nosp, (-1, -1, -1), T, 1234)
    got = tokens[:, n_init].long()
    masked = row[0, :V].double().clone()
    masked[sup.bool()] = -np.inf
    p = torch.softmax(masked / T, -1)
    assert not sup.bool()[got].any()
    top = torch.topk(p, 6).indices
    emp = torch.stack([(got == i).double().mean() for i in top])
    sigma = torch.sqrt(p[top] * (1 - p[top]) / R)
    assert ((emp - p[top]).abs() <= 5 * sigma + 5e-4).all(), (emp, p[top])
    want_lp = torch.log_softmax(masked, -1)[got]
    assert (slp.double() - want_lp).abs().max().item() <= 1e-3
    # deterministic in (seed, row, position); a different seed gives a different draw
    tok2 = tokens.clone(); tok2[:, n_init:] = 0
    nv.sample_greedy(logits, V, sup, None, tok2, state, slp, nsp, eot, nosp, (-1, -1, -1), T, 1234)
    assert torch.equal(tok2[:, n_init], tokens[:, n_init])
    nv.sample_greedy(logits, V, sup, None, tok2, state, slp, nsp, eot, nosp, (-1, -1, -1), T, 99)
    assert not torch.equal(tok2[:, n_init], tokens[:, n_init])


def test_kv_gather_rows(nv):
    src = _randn(8, 10, 128, dtype=torch.bfloat16, seed=1)
    dst = torch.zeros_like(src)
    index = torch.tensor([3, 3, 0, 7, 1, 1, 1, 2], dtype=torch.int32, device="cuda")
    nv.kv_gather_rows(src, dst, index, 8, 10 * 128 * 2, 4 * 128 * 2)
    assert torch.equal(dst[:, :4], src[index.long(), :4]) and dst[:, 4:].abs().max().item() == 0


def test_errors_are_reported_not_swallowed(nv):
    a = _randn(8, 60, dtype=torch.bfloat16)        # K * 2 bytes = 120: not a multiple of 16 -> TMA cannot map it
    w = _randn(8, 60, dtype=torch.bfloat16)
    with pytest.raises(nv.WfError):
        nv.linear(a, w, torch.empty((8, 8), dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(nv.WfError):
        nv.logmel(torch.zeros(1, 100, device="cuda"), 80, 0)


# ----------------------------------------------------------------------------- absorbed ("latent") cross-attention
def _latent_reference(q, src, wk, wv, bv, H):
    """model.py:82-108 restated in fp32: K = src Wk^T, V = src Wv^T + bv, o_h = softmax(q_h K_h^T / 8) V_h."""
    B, T, d = src.shape
    q, src, wk, wv = q.float(), src.float(), wk.float(), wv.float()
    k = (src @ wk.T).view(B, T, H, 64).permute(0, 2, 1, 3)
    v = (src @ wv.T + bv).view(B, T, H, 64).permute(0, 2, 1, 3)
    qh = q.view(B, H, 1, 64)
    w = torch.softmax((qh @ k.transpose(-1, -2)) * 0.125, dim=-1)
    return (w @ v).reshape(B, H * 64)


@pytest.mark.parametrize("H,B,T", [(20, 3, 1500), (20, 2, 750), (16, 2, 300), (12, 2, 130), (8, 1, 128), (4, 2, 40),
                                   (20, 5, 1)])
def test_latent_cross_attention_matches_kv_attention(nv, H, B, T):
    from helpers import rel_l2
    d = 64 * H
    bf = torch.bfloat16
    q = _randn(B, d, dtype=bf, seed=1)
    src = _randn(B, T, d, dtype=bf, seed=2)
    wk = _randn(d, d, dtype=bf, seed=3, scale=2.0 / math.sqrt(d))   # scores with a spread of a few units
    wv = _randn(d, d, dtype=bf, seed=4, scale=1.0 / math.sqrt(d))
    bv = _randn(d, seed=5, scale=0.1)
    qp = torch.empty(B, H, d, dtype=bf, device="cuda")
    ctx = torch.full((B, H, d), float("nan"), dtype=bf, device="cuda")
    out = torch.full((B, d), float("nan"), dtype=bf, device="cuda")
    nv.latent_query(q, wk.t().contiguous(), qp, H)
    nv.latent_attention(qp, src, ctx, H)
    nv.latent_value(ctx, wv, bv, out, H)
    torch.cuda.synchronize()
    # stage 1: q' = Wk_h^T q_h
    qp_ref = torch.einsum("bhj,hjn->bhn", q.float().view(B, H, 64), wk.float().view(H, 64, d))
    assert rel_l2(qp, qp_ref) < 6e-3
    # stage 2 (from the kernel's own bf16 q'): softmax over keys, context = weighted sum of source rows
    sc = torch.einsum("bhn,btn->bht", qp.float(), src.float()) * 0.125
    ctx_ref = torch.einsum("bht,btn->bhn", torch.softmax(sc, dim=-1), src.float())
    assert rel_l2(ctx, ctx_ref) < 8e-3
    # stage 3 (from the kernel's own bf16 context)
    out3 = torch.einsum("bhn,hjn->bhj", ctx.float(), wv.float().view(H, 64, d)).reshape(B, d) + bv
    assert rel_l2(out, out3) < 6e-3
    # end to end against attention over projected K / V (what the reference computes); bf16 tolerance: the K / V
    # rounding of the cached path is replaced by the rounding of q' and of the context
    assert rel_l2(out, _latent_reference(q, src, wk, wv, bv, H)) < 1.5e-2


@pytest.mark.parametrize("H,B,T", [(20, 128, 1500), (20, 128, 750), (16, 100, 333), (20, 75, 64), (12, 80, 130),
                                   (20, 5, 200)])
def test_latent_split_form_matches_single_part(nv, H, B, T):
    """Full batches run the persistent pair kernel: clusters take equal tile ranges, a clip cut at a range border is
    left as two separately normalised contexts + (maximum, row sum) per head, and latent_value blends them.  The blend
    must equal attention over projected K / V, every clip must have a first part, and a second part exists exactly
    where the partition of B * ceil(T / 64) tiles over the SM pairs cuts a clip."""
    from helpers import rel_l2
    if not nv.latent_split_supported(H):
        pytest.skip("no pair kernel for this width")
    d = 64 * H
    bf = torch.bfloat16
    q = _randn(B, d, dtype=bf, seed=21)
    src = _randn(B, T, d, dtype=bf, seed=22)
    wk = _randn(d, d, dtype=bf, seed=23, scale=2.0 / math.sqrt(d))
    wv = _randn(d, d, dtype=bf, seed=24, scale=1.0 / math.sqrt(d))
    bv = _randn(d, seed=25, scale=0.1)
    qp = torch.empty(B, H, d, dtype=bf, device="cuda")
    ctx = torch.zeros(2, B, H, d, dtype=bf, device="cuda")
    ml = torch.full((2, B, 32, 2), float("nan"), device="cuda")
    out = torch.full((B, d), float("nan"), dtype=bf, device="cuda")
    nv.latent_query(q, wk.t().contiguous(), qp, H)
    nv.latent_attention(qp, src, ctx, H, ml=ml)
    nv.latent_value(ctx, wv, bv, out, H, ml=ml)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    l = ml[..., 1][:, :, :H]
    assert torch.isfinite(l).all() and (l[0] > 0).all(), "every clip has a first part"
    # where the second part exists: a clip is cut iff a cluster range border falls strictly inside it
    n_tiles = (T + 63) // 64
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clusters = sms // 2 if B > sms // 2 else B
    borders = {B * n_tiles * c // clusters for c in range(1, clusters)}
    cut = torch.tensor([any(b * n_tiles < g < (b + 1) * n_tiles for g in borders) for b in range(B)], device="cuda")
    assert torch.equal((l[1] > 0).all(dim=1), cut) and torch.equal((l[1] > 0).any(dim=1), cut)
    for lo in range(0, B, 16):
        ref = _latent_reference(q[lo:lo + 16], src[lo:lo + 16], wk, wv, bv, H)
        assert rel_l2(out[lo:lo + 16], ref) < 1.5e-2
    # against the single-part form of the same kernel family (same rounding points up to the blend)
    ctx1 = torch.empty(B, H, d, dtype=bf, device="cuda")
    out1 = torch.empty(B, d, dtype=bf, device="cuda")
    nv.latent_attention(qp, src, ctx1, H)
    nv.latent_value(ctx1, wv, bv, out1, H)
    torch.cuda.synchronize()
    assert rel_l2(out, out1) < 6e-3
    whole = ~cut
    assert torch.equal(ctx[0][whole], ctx1[whole]), "uncut clips: same tiles in the same order, bit for bit"


def test_latent_cross_attention_headline_size(nv):
    """BASELINE config 4 shapes (128 clips x 1500 encoder rows, 20 heads): the latent path against attention over
    projected K / V, and two properties that do not need a reference: the result does not depend on which clips share
    the launch (clip 5 alone == clip 5 inside the batch, bit for bit) and the context of every head is a convex
    combination of source rows (bounded by the row-wise min / max of its clip)."""
    from helpers import rel_l2
    H, B, T = 20, 128, 1500
    d = 64 * H
    bf = torch.bfloat16
    q = _randn(B, d, dtype=bf, seed=11)
    src = _randn(B, T, d, dtype=bf, seed=12)
    wk = _randn(d, d, dtype=bf, seed=13, scale=2.0 / math.sqrt(d))
    wv = _randn(d, d, dtype=bf, seed=14, scale=1.0 / math.sqrt(d))
    bv = _randn(d, seed=15, scale=0.1)
    qp = torch.empty(B, H, d, dtype=bf, device="cuda")
    ctx = torch.full((B, H, d), float("nan"), dtype=bf, device="cuda")
    out = torch.full((B, d), float("nan"), dtype=bf, device="cuda")
    nv.latent_query(q, wk.t().contiguous(), qp, H)
    nv.latent_attention(qp, src, ctx, H)
    nv.latent_value(ctx, wv, bv, out, H)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(ctx.float()).all()
    for lo in range(0, B, 16):   # reference in slices: fp32 K and V of 16 clips are 2 x 123 MB
        ref = _latent_reference(q[lo:lo + 16], src[lo:lo + 16], wk, wv, bv, H)
        assert rel_l2(out[lo:lo + 16], ref) < 1.5e-2
    lo_, hi_ = src.float().amin(dim=1), src.float().amax(dim=1)          # [B, d]
    c = ctx.float()
    assert (c >= lo_[:, None, :] - 2e-2).all() and (c <= hi_[:, None, :] + 2e-2).all()
    one = torch.empty(1, H, d, dtype=bf, device="cuda")
    nv.latent_attention(qp[5:6].contiguous(), src[5:6].contiguous(), one, H)
    torch.cuda.synchronize()
    assert torch.equal(one[0], ctx[5])
