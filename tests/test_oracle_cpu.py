"""CPU suite: the oracle against the committed golden fixtures (reference outputs), host logic,
and the C-ABI surface of libwf.so.  No GPU needed."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, SAMPLE_IDX, TINY, build_model, load_decode_golden, oracle_sd, spec_from_json
from oracle import decode as odec
from oracle import mel as omel
from oracle import model as om

ROOT = os.path.dirname(GOLDEN.rstrip("/").rsplit("/tests", 1)[0] + "/x")


def _pcm(n_clips, seed=1234):
    from whisper._synthetic import synthetic_pcm
    return synthetic_pcm(n_clips, seed=seed).numpy()


# ----------------------------------------------------------------------------- mel
def test_mel_filterbank_sha256_matches_reference_asset():
    # sha256 prefixes of the raw fp32 bytes of the reference's assets/mel_filters.npz
    assert hashlib.sha256(omel.mel_filterbank(80).tobytes()).hexdigest().startswith("4f2701b1d287d74a")
    assert hashlib.sha256(omel.mel_filterbank(128).tobytes()).hexdigest().startswith("2a5f9822897750e0")
    from whisper.audio import _slaney_mel_filterbank
    assert np.array_equal(_slaney_mel_filterbank(80), omel.mel_filterbank(80))
    assert np.array_equal(_slaney_mel_filterbank(128), omel.mel_filterbank(128))


@pytest.mark.parametrize("n_mels", [80, 128])
def test_oracle_mel_matches_reference_golden(n_mels):
    g = np.load(os.path.join(GOLDEN, "mel.npz"))
    pcm = _pcm(2)
    m = omel.log_mel_spectrogram(pcm[0], n_mels)
    assert m.shape == (n_mels, 3000)
    idx = SAMPLE_IDX[SAMPLE_IDX < m.size]
    assert np.abs(m.reshape(-1)[idx] - g[f"gauss{n_mels}_samples"]).max() <= 5e-5
    c = omel.log_mel_spectrogram(omel.chirp_kat(), n_mels)
    assert np.abs(c.reshape(-1)[idx] - g[f"chirp{n_mels}_samples"]).max() <= 1.5e-4
    lo, hi, mean = g[f"chirp{n_mels}_stats"]
    assert abs(c.min() - lo) < 1e-4 and abs(c.max() - hi) < 1e-4 and abs(c.mean() - mean) < 1e-4
    assert abs((c.max() - c.min()) - 2.0) < 1e-6  # the max-8 clamp binds on the chirp


def test_chirp_known_answers_from_survey():
    # values captured from the live reference at survey time
    m = omel.log_mel_spectrogram(omel.chirp_kat(), 80)
    for (i, j), v in {(0, 0): 1.393853, (10, 100): 1.040327, (79, 2999): 0.439978, (40, 1500): -0.547307}.items():
        assert abs(m[i, j] - v) < 2e-4, (i, j, m[i, j], v)
    m = omel.log_mel_spectrogram(omel.chirp_kat(), 128)
    for (i, j), v in {(0, 0): 1.318094, (10, 100): 0.744588, (127, 2999): 0.386699}.items():
        assert abs(m[i, j] - v) < 2e-4, (i, j, m[i, j], v)


def test_oracle_mel_batched_global_max_and_padding():
    g = np.load(os.path.join(GOLDEN, "mel.npz"))
    pcm = _pcm(2)
    two = np.stack([pcm[0], pcm[1] * 1e-3]).astype(np.float32)
    b = omel.log_mel_spectrogram(two, 80)
    half = b.size // 2
    idx = SAMPLE_IDX[SAMPLE_IDX < half]
    got = np.concatenate([b.reshape(-1)[idx], b.reshape(-1)[half + idx]])
    assert np.abs(got - g["batch2_global_samples"]).max() <= 1e-5
    # per-clip max differs from the reference's global max when the levels differ
    pc = omel.log_mel_spectrogram(two, 80, per_clip_max=True)
    assert np.abs(pc[1] - b[1]).max() > 0.1 and np.abs(pc[0] - b[0]).max() < 1e-6
    sp = omel.log_mel_spectrogram(pcm[0][:16000], 80, padding=4800)
    assert sp.shape == (80, 130) and np.abs(sp - g["short_pad_full"]).max() <= 1e-5


def test_pad_or_trim_matches_oracle():
    import whisper
    x = np.arange(10, dtype=np.float32)
    for n in (4, 10, 16):
        assert np.array_equal(whisper.pad_or_trim(x, n), omel.pad_or_trim(x, n))
        assert np.array_equal(whisper.pad_or_trim(torch.from_numpy(x), n).numpy(), omel.pad_or_trim(x, n))
    y = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    assert np.array_equal(whisper.pad_or_trim(y, 5, axis=1), omel.pad_or_trim(y, 5, axis=1))
    assert np.array_equal(whisper.pad_or_trim(torch.from_numpy(y), 2, axis=1).numpy(), omel.pad_or_trim(y, 2, axis=1))
    assert np.array_equal(whisper.pad_or_trim(torch.from_numpy(y), 6, axis=0).numpy(), omel.pad_or_trim(y, 6, axis=0))


# ----------------------------------------------------------------------------- network
def test_oracle_network_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "net_tiny_av.npz"))
    from whisper._synthetic import synthetic_features
    model = build_model(gated=True)
    sd = oracle_sd(model)
    dims = om.Dims(**TINY)
    mel = torch.from_numpy(omel.log_mel_spectrogram(_pcm(1), 80))
    feat = synthetic_features(1, n_frames=100, dim=1024, seed=4321)
    toks = torch.from_numpy(g["tokens"])
    with torch.no_grad():
        xa = om.encoder_forward(sd, dims, mel)
        lg = om.decoder_forward(sd, dims, toks, xa, xt_list=[feat])
    assert np.abs(xa.reshape(-1)[::4001].numpy() - g["xa_samples"]).max() < 5e-4
    assert np.abs(xa[0, 0].numpy() - g["xa_row0"]).max() < 5e-4
    assert np.abs(lg.reshape(-1)[::1009].numpy() - g["logits_samples"]).max() < 2e-3
    assert torch.topk(lg[0, -1], 8).indices.tolist() == g["logits_last_top"].tolist()


def test_oracle_rejects_too_many_feature_tensors():
    model = build_model(gated=True)
    sd, dims = oracle_sd(model), om.Dims(**TINY)
    xa = torch.zeros(1, 4, 384)
    f = torch.zeros(1, 3, 1024)
    with pytest.raises(ValueError):
        om.decoder_forward(sd, dims, torch.tensor([[50258]]), xa, xt_list=[f, f])


# ----------------------------------------------------------------------------- decode loops
def test_oracle_greedy_tokens_match_reference_golden():
    gold = load_decode_golden()["cases"]["greedy_audio_only"]
    spec = spec_from_json(gold["spec"])
    spec.sample_len = 12  # the first 12 of the reference's 64 greedy tokens (keeps the CPU suite short)
    model = build_model(gated=False)
    mel = torch.from_numpy(omel.log_mel_spectrogram(_pcm(1), 80))
    res = odec.decode(oracle_sd(model), om.Dims(**TINY), spec, mel)
    assert res[0].tokens == gold["tokens"][:12]
    assert abs(res[0].no_speech_prob - gold["no_speech_prob"]) < 1e-6


def test_oracle_timestamp_rules_match_reference_golden():
    gold = load_decode_golden()["cases"]["greedy_timestamps"]
    spec = spec_from_json(gold["spec"])
    spec.sample_len = 8
    model = build_model(gated=False)
    mel = torch.from_numpy(omel.log_mel_spectrogram(_pcm(1), 80))
    res = odec.decode(oracle_sd(model), om.Dims(**TINY), spec, mel)
    assert res[0].tokens == gold["tokens"][:8]
    assert res[0].tokens[0] >= spec.timestamp_begin  # first sampled token must be a timestamp


def test_oracle_beam_matches_reference_golden():
    gold = load_decode_golden()["cases"]["beam3_audio_only"]
    spec = spec_from_json(gold["spec"])
    model = build_model(gated=False)
    mel = torch.from_numpy(omel.log_mel_spectrogram(_pcm(1), 80))
    res = odec.decode(oracle_sd(model), om.Dims(**TINY), spec, mel)
    assert res[0].tokens == gold["tokens"]
    assert abs(res[0].avg_logprob - gold["avg_logprob"]) < 1e-5


def test_oracle_detect_language_matches_reference_golden():
    """detect_language restatement (decoding.py:18-77) against the reference's output on 2 clips."""
    from whisper.tokenizer import get_tokenizer
    gold = load_decode_golden()["cases"]["detect_language"]
    model = build_model(gated=False)
    sd, dims = oracle_sd(model), om.Dims(**TINY)
    mel = torch.from_numpy(np.stack([omel.log_mel_spectrogram(p, 80) for p in _pcm(2)]))
    tk = get_tokenizer(True, num_languages=99)
    with torch.no_grad():
        best, probs = odec.detect_language(sd, dims, om.encoder_forward(sd, dims, mel), tk.sot,
                                           list(tk.all_language_tokens))
    assert best == gold["language_tokens"]
    codes = list(tk.all_language_codes)
    assert [codes[int(i)] for i in probs.argmax(-1)] == gold["languages"]
    for i in range(2):
        assert abs(float(probs[i].max()) - gold["top_prob"][i]) < 1e-6
        assert abs(float(probs[i, codes.index("en")]) - gold["probs_en"][i]) < 1e-7


def test_oracle_multi_feature_greedy_matches_reference_golden():
    """xt_list with 3 tensors of different lengths / widths (model.py:171-199), greedy tokens of the reference loop."""
    from whisper._synthetic import synthetic_features
    gold = load_decode_golden()["cases"]["greedy_multi3"]
    spec = spec_from_json(gold["spec"])
    spec.sample_len = 6
    model = build_model(gated=True, num_langs=3)
    mel = torch.from_numpy(np.stack([omel.log_mel_spectrogram(p, 80) for p in _pcm(2)]))
    feats = [synthetic_features(2, n_frames=s[0], dim=s[1], seed=seed)
             for s, seed in zip(gold["feat_shapes"], gold["feat_seeds"])]
    res = odec.decode(oracle_sd(model), om.Dims(**TINY), spec, mel, feats)
    assert [r.tokens for r in res] == [t[:6] for t in gold["tokens"]]
    g = np.load(f"{GOLDEN}/net_tiny_multi.npz")
    sd, dims = oracle_sd(model), om.Dims(**TINY)
    with torch.no_grad():
        lg = om.decoder_forward(sd, dims, torch.from_numpy(g["tokens"]), om.encoder_forward(sd, dims, mel), xt_list=feats)
    assert np.abs(lg.reshape(-1)[::1009].numpy() - g["logits3_samples"]).max() < 2e-3


# ----------------------------------------------------------------------------- host logic of the drop-in
def test_tokenizer_special_ids_and_suppress_set():
    from whisper.tokenizer import get_tokenizer
    tk = get_tokenizer(True, num_languages=99, language="en", task="transcribe")
    assert (tk.eot, tk.sot, tk.translate, tk.transcribe, tk.sot_lm, tk.sot_prev, tk.no_speech, tk.no_timestamps,
            tk.timestamp_begin) == (50257, 50258, 50358, 50359, 50360, 50361, 50362, 50363, 50364)
    assert tk.sot_sequence_including_notimestamps == (50258, 50259, 50359, 50363)
    assert tk.encode(" ") == [220] and len(tk.non_speech_tokens) == 82
    assert tk.decode(tk.encode("hello world")) == "hello world"
    gold = load_decode_golden()["cases"]["greedy_audio_only"]
    assert tk.decode(gold["tokens"]).strip() == gold["text"]
    with pytest.raises(ValueError):
        get_tokenizer(True, language="klingon")
    en = get_tokenizer(False)
    assert en.sot_sequence == (50257,) and en.eot == 50256


def test_decoding_task_options_match_reference_spec():
    import whisper
    from whisper.decoding import DecodingTask
    gold = load_decode_golden()["cases"]
    model = build_model(gated=False)
    task = DecodingTask(model, whisper.DecodingOptions(language="en", without_timestamps=True, sample_len=64, fp16=False))
    spec = gold["greedy_audio_only"]["spec"]
    assert list(task.initial_tokens) == spec["initial_tokens"]
    assert list(task._get_suppress_tokens()) == spec["suppress_tokens"]
    assert task.sample_begin == 4 and task.sot_index == 0 and task.ts_params == (-1, -1, -1)
    task = DecodingTask(model, whisper.DecodingOptions(language="en", sample_len=24, fp16=False))
    spec = gold["greedy_timestamps"]["spec"]
    assert list(task.initial_tokens) == spec["initial_tokens"]
    assert task.ts_params == (spec["timestamp_begin"], spec["no_timestamps"], spec["max_initial_timestamp_index"])
    # prompt / prefix splice (reference decoding.py:591-617)
    task = DecodingTask(model, whisper.DecodingOptions(language="en", prompt=[11, 12, 13], prefix=[21, 22],
                                                       without_timestamps=True))
    assert task.initial_tokens == (50361, 11, 12, 13, 50258, 50259, 50359, 50363, 21, 22) and task.sot_index == 4
    for bad in (dict(beam_size=2, best_of=2), dict(best_of=2), dict(patience=1.0), dict(length_penalty=2.0)):
        with pytest.raises(ValueError):
            DecodingTask(model, whisper.DecodingOptions(**bad))


def test_state_dict_names_match_reference_layout():
    model = build_model(gated=True)
    keys = set(model.state_dict().keys())
    for k in ("encoder.conv1.weight", "encoder.positional_embedding", "encoder.blocks.3.attn.key.weight",
              "encoder.ln_post.bias", "decoder.token_embedding.weight", "decoder.positional_embedding",
              "decoder.xt_projection.weight", "decoder.blocks.0.cross_attn.out.bias",
              "decoder.blocks.2.gated_x_attn_layers.0.attn_gate", "decoder.blocks.2.gated_x_attn_layers.0.attn_ln.weight",
              "decoder.blocks.1.ff_ln.weight", "decoder.blocks.1.ff.2.bias", "decoder.blocks.1.ff_gate"):
        assert k in keys, k
    assert "encoder.blocks.0.attn.key.bias" not in keys and "decoder.mask" not in keys
    assert model.is_multilingual and model.num_languages == 99


def test_product_has_no_cpu_fallback():
    import whisper
    model = build_model(gated=False)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-GPU failure mode is not observable")
    with pytest.raises(RuntimeError):
        whisper.log_mel_spectrogram(torch.zeros(16000))
    with pytest.raises(RuntimeError):
        model.encoder(torch.zeros(1, 80, 3000))
    with pytest.raises(RuntimeError):
        whisper.decode(model, torch.zeros(80, 3000), whisper.DecodingOptions(language="en"))


def test_product_never_imports_oracle():
    pkg = os.path.join(os.path.dirname(GOLDEN), "..", "whisper-flamingo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
                assert not re.search(r"^\s*(from|import)\s+baseline", src, re.M), f"{f} imports the reference arm"
                assert "ref_whisper" not in src and "/root/reference" not in src, f"{f} reaches for the reference"


def test_reference_arm_is_only_used_by_bench_and_is_git_ignored():
    """baseline/reference_arm.py drives the unmodified reference for bench.py; the staged copy stays out of history."""
    repo = os.path.join(os.path.dirname(GOLDEN), "..")
    ignore = open(os.path.join(repo, ".gitignore")).read()
    assert "baseline/_ref/" in ignore
    gpurunignore = os.path.join(repo, ".gpurunignore")
    assert not os.path.exists(gpurunignore) or "baseline/_ref" not in open(gpurunignore).read()
    from baseline import reference_arm as ra
    assert ra.CANDIDATES[0].endswith("baseline/_ref/whisper") and ra.CANDIDATES[1] == "/root/reference/whisper"


# ----------------------------------------------------------------------------- C ABI surface
def test_libwf_exports_every_declared_symbol():
    from whisper import _native
    if not os.path.exists(_native.lib_path()):
        pytest.skip("libwf.so not built (run python __graft_entry__.py build)")
    header = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "wf.h")).read()
    declared = set(re.findall(r"\b(wf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed from include/wf.h"
    lib = ctypes.CDLL(_native.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"libwf.so does not export {name}"
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    assert _native.load().wf_version() >= 100


def test_load_model_accepts_fork_lightning_and_upstream_checkpoints(tmp_path):
    """checkpoint ingestion.  The same weights saved (a) in the fork's own format, (b) as a Lightning
    checkpoint with the `model.` prefix and no dims, (c) with upstream Whisper-Flamingo key names must all load into
    identical parameters (reference whisper/__init__.py:152-159, whisper-flamingo_kloka_crawled.py:172-183)."""
    import whisper
    from helpers import TINY, build_model
    dims = dict(TINY, n_audio_layer=1, n_text_layer=2)
    model = build_model(gated=True, device="cpu", dims=dims)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    a = tmp_path / "fork.pt"
    torch.save({"dims": dims, "model_state_dict": sd}, a)

    def upstream(key):
        key = key.replace("gated_x_attn_layers.0.attn_ln.", "gated_x_attn_ln.")
        key = key.replace("gated_x_attn_layers.0.attn_gate", "attn_gate")
        return key.replace("gated_x_attn_layers.0.attn.", "gated_x_attn.")

    b = tmp_path / "lightning.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}, "epoch": 3}, b)
    c = tmp_path / "upstream.ckpt"
    up = {"model." + upstream(k): v for k, v in sd.items()}
    up["model.encoder.video_projection_scalar"] = torch.ones(1)  # upstream-only key: ignored (strict=False)
    assert any(".gated_x_attn.query.weight" in k for k in up) and not any("gated_x_attn_layers" in k for k in up)
    torch.save({"state_dict": up}, c)
    for path in (a, b, c):
        got = whisper.load_model(str(path), device="cpu", add_gated_x_attn=1, bert_dim=1024, num_langs=1)
        assert got.dims == whisper.ModelDimensions(**dims)
        gsd = got.state_dict()
        assert gsd.keys() == sd.keys()
        assert all(torch.equal(gsd[k], sd[k]) for k in sd), path
    with pytest.raises(RuntimeError):
        whisper.load_model("no-such-model")


# ----------------------------------------------------------------------------- latent cross-attention (host side)
def test_absorbed_projection_identity_against_oracle_mha():
    """The algebra csrc/latent.cu relies on, in fp64 against the oracle's restatement of model.py:82-108:
    with q'_h = Wk_h^T q_h, c_h = softmax(xa q'_h / 8)^T xa and o_h = Wv_h c_h + bv_h the one-token cross-attention
    equals attention over K = xa Wk^T, V = xa Wv^T + bv (the softmax weights sum to one, so the value bias passes)."""
    g = torch.Generator().manual_seed(7)
    B, T, H = 2, 37, 3
    d = 64 * H
    r = lambda *shape: torch.randn(*shape, generator=g, dtype=torch.float64)
    sd = {"a.query.weight": r(d, d) / d ** 0.5, "a.query.bias": r(d) * 0.1, "a.key.weight": r(d, d) / d ** 0.5,
          "a.value.weight": r(d, d) / d ** 0.5, "a.value.bias": r(d) * 0.1, "a.out.weight": r(d, d) / d ** 0.5,
          "a.out.bias": r(d) * 0.1}
    x, xa = r(B, 1, d), r(B, T, d)
    want = om.mha(sd, "a", H, x, xa)
    q = x[:, 0] @ sd["a.query.weight"].T + sd["a.query.bias"]                          # [B, d]
    qp = torch.einsum("bhj,hjn->bhn", q.view(B, H, 64), sd["a.key.weight"].view(H, 64, d))    # q'_h = Wk_h^T q_h
    w = torch.softmax(torch.einsum("bhn,btn->bht", qp, xa) * 0.125, dim=-1)
    c = torch.einsum("bht,btn->bhn", w, xa)                                            # context in source space
    o = torch.einsum("bhn,hjn->bhj", c, sd["a.value.weight"].view(H, 64, d)).reshape(B, d) + sd["a.value.bias"]
    got = (o @ sd["a.out.weight"].T + sd["a.out.bias"])[:, None]
    # the oracle keeps the reference's fp32 softmax (model.py:104), everything else here is fp64
    assert torch.allclose(got, want, rtol=0, atol=5e-7), float((got - want).abs().max())


def test_latent_path_gate(monkeypatch):
    """whisper/_engine.py:latent_cross_enabled - greedy rows only, head_dim 64, at most 24 heads, default from 112 rows."""
    from whisper._engine import latent_cross_enabled as on
    monkeypatch.delenv("WF_LATENT", raising=False)
    assert on(128, 1, 20, 1280) and on(112, 1, 16, 1024) and not on(111, 1, 20, 1280)
    assert not on(128, 5, 16, 1024)            # beams share a cache: cached-K/V multi-query kernel
    assert not on(128, 1, 40, 2560) and not on(128, 1, 20, 1024) and not on(128, 1, 5, 320)
    monkeypatch.setenv("WF_LATENT", "1")
    assert on(2, 1, 6, 384) and not on(2, 5, 6, 384)
    monkeypatch.setenv("WF_LATENT", "0")
    assert not on(128, 1, 20, 1280)


# ----------------------------------------------------------------------------- load_audio (reference audio.py:26-63)
def test_load_audio_reads_wave_files_without_ffmpeg(tmp_path, monkeypatch):
    """Hosts without ffmpeg (this image) read RIFF/WAVE natively: 16-bit PCM at 16 kHz is exactly what the reference's
    ffmpeg command would emit (int16 / 32768); other rates go through the windowed-sinc resampler, checked against the
    analytic signal; stereo is down-mixed; a non-WAVE file is refused with the reference's error type."""
    import struct
    import whisper
    from whisper.audio import load_audio
    import shutil
    if shutil.which("ffmpeg"):
        monkeypatch.setenv("PATH", "")           # exercise the fall-back everywhere

    def write_wav(path, x, rate, channels=1, bits=16):
        if bits == 16:
            data = (np.clip(x, -1, 1) * 32767).astype("<i2").tobytes()
        else:
            data = x.astype("<f4").tobytes()
        fmt = struct.pack("<HHIIHH", 1 if bits == 16 else 3, channels, rate, rate * channels * bits // 8,
                          channels * bits // 8, bits)
        with open(path, "wb") as fh:
            fh.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt)
            fh.write(b"LIST" + struct.pack("<I", 4) + b"abcd")            # an extra chunk to skip
            fh.write(b"data" + struct.pack("<I", len(data)) + data)

    t = np.arange(16000) / 16000.0
    x = 0.5 * np.sin(2 * np.pi * 440 * t)
    write_wav(tmp_path / "a.wav", x, 16000)
    got = load_audio(str(tmp_path / "a.wav"))
    assert got.dtype == np.float32 and got.shape == (16000,)
    assert np.array_equal(got, (x * 32767).astype("<i2").astype(np.float32) / 32768.0)
    # 44.1 kHz stereo float -> 16 kHz mono
    t2 = np.arange(44100) / 44100.0
    st = np.stack([0.5 * np.sin(2 * np.pi * 440 * t2), 0.5 * np.sin(2 * np.pi * 440 * t2)], axis=1).reshape(-1)
    write_wav(tmp_path / "b.wav", st, 44100, channels=2, bits=32)
    got = load_audio(str(tmp_path / "b.wav"))
    assert got.shape == (16000,)
    assert np.abs(got[200:-200] - x[200:-200]).max() < 2e-3
    assert whisper.load_audio is load_audio
    (tmp_path / "c.bin").write_bytes(b"not audio at all")
    with pytest.raises(RuntimeError, match="Failed to load audio"):
        load_audio(str(tmp_path / "c.bin"))
