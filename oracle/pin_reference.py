"""Pin the oracle against the LIVE reference and write the golden fixtures (TEST INFRASTRUCTURE).

Runs only in the authoring container, where the reference is mounted read-only at /root/reference
(``PYTHONDONTWRITEBYTECODE=1 python oracle/pin_reference.py``).  It

1. imports the unmodified reference package under the alias ``ref_whisper``;
2. checks every oracle restatement (mel, encoder, decoder, greedy / beam / timestamp decode loops)
   against the reference on seeded inputs - any mismatch aborts;
3. writes ``tests/golden/*.npz|json``: reference outputs on those inputs.  The fixtures travel to
   the GPU box (the reference does not); ``tests/`` compares both the oracle and the CUDA path to them.

The reference has no tests or golden vectors of its own, so these fixtures - outputs
of the reference itself on torch 2.11.0 CPU - are what "parity" is pinned to.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import decode as odec  # noqa: E402
from oracle import mel as omel  # noqa: E402
from oracle import model as om  # noqa: E402


def load_by_path(alias: str, path: str, package: bool = False):
    kw = {"submodule_search_locations": [os.path.dirname(path)]} if package else {}
    spec = importlib.util.spec_from_file_location(alias, path, **kw)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


ref = load_by_path("ref_whisper", os.path.join(REF, "whisper", "__init__.py"), package=True)
synth = load_by_path("wf_synthetic", os.path.join(ROOT, "whisper-flamingo_b200", "whisper", "_synthetic.py"))
from ref_whisper.decoding import DecodingOptions, DecodingTask  # noqa: E402
from ref_whisper.model import ModelDimensions, Whisper  # noqa: E402

TINY = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4, n_vocab=51865,
            n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)
SAMPLE_IDX = np.arange(0, 80 * 3000, 977)  # fixed subsample grid for large tensors


def check(name, a, b, tol):
    err = float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
    status = "ok" if err <= tol else "MISMATCH"
    print(f"  [{status}] {name}: max abs diff {err:.3e} (tol {tol:.1e})")
    if err > tol:
        raise SystemExit(f"oracle disagrees with the reference on {name}")
    return err


def spec_from_task(task: DecodingTask, options: DecodingOptions) -> odec.DecodeSpec:
    tk = task.tokenizer
    mi = None
    if not options.without_timestamps:
        mi = task.logit_filters[-1].max_initial_timestamp_index
    return odec.DecodeSpec(
        initial_tokens=tuple(task.initial_tokens), eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech,
        suppress_tokens=tuple(task._get_suppress_tokens()) if options.suppress_tokens else (),
        blank_tokens=tuple(tk.encode(" ") + [tk.eot]) if options.suppress_blank else (),
        sample_len=task.sample_len, n_ctx=task.n_ctx, beam_size=options.beam_size, patience=options.patience,
        length_penalty=options.length_penalty, without_timestamps=options.without_timestamps,
        timestamp_begin=tk.timestamp_begin, no_timestamps=tk.no_timestamps, max_initial_timestamp_index=mi)


def spec_to_json(spec: odec.DecodeSpec) -> dict:
    d = dict(spec.__dict__)
    d["initial_tokens"] = list(spec.initial_tokens)
    d["suppress_tokens"] = list(spec.suppress_tokens)
    d["blank_tokens"] = list(spec.blank_tokens)
    return d


@torch.no_grad()
def reference_av_greedy(model, task: DecodingTask, mel, feat):
    """the reference's own loop pieces + the missing ``xt_list`` argument."""
    task.decoder.reset()
    xa = model.encoder(mel)
    tokens = torch.tensor([task.initial_tokens]).repeat(mel.shape[0], 1)
    sum_lp = torch.zeros(tokens.shape[0])
    no_speech = None
    for i in range(task.sample_len):
        logits = model.decoder(tokens, xa, xt_list=[feat])
        if i == 0:
            no_speech = logits[:, task.sot_index].float().softmax(-1)[:, task.tokenizer.no_speech].tolist()
        logits = logits[:, -1]
        for f in task.logit_filters:
            f.apply(logits, tokens)
        tokens, done = task.decoder.update(tokens, logits, sum_lp)
        if done or tokens.shape[-1] > task.n_ctx:
            break
    eot = task.tokenizer.eot
    out = []
    for row, lp in zip(tokens.tolist(), sum_lp.tolist()):
        row = row[task.sample_begin:] + [eot]
        toks = row[: row.index(eot)]
        out.append((toks, lp / (len(toks) + 1)))
    return out, no_speech


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    meta = {"torch": torch.__version__, "reference": "jerryyang1231/whisper-flamingo @ /root/reference"}

    # ------------------------------------------------------------------ 1. log-mel
    print("== log-mel")
    filt = np.load(os.path.join(REF, "whisper", "assets", "mel_filters.npz"))
    for n in (80, 128):
        assert np.array_equal(omel.mel_filterbank(n), filt[f"mel_{n}"]), f"mel_{n} filterbank not bit-exact"
    print("  [ok] regenerated mel filterbanks are bit-exact (80, 128)")
    pcm = synth.synthetic_pcm(2, seed=1234).numpy()
    chirp = omel.chirp_kat()
    mel_gold = {}
    for n in (80, 128):
        r = ref.log_mel_spectrogram(torch.from_numpy(pcm[0]), n_mels=n).numpy()
        check(f"mel{n} gaussian (oracle fp64 vs ref)", omel.log_mel_spectrogram(pcm[0], n), r, 5e-5)
        check(f"mel{n} gaussian (oracle fp32 vs ref)", omel.log_mel_spectrogram(pcm[0], n, dtype=np.float32), r, 5e-5)
        rc = ref.log_mel_spectrogram(torch.from_numpy(chirp), n_mels=n).numpy()
        check(f"mel{n} chirp KAT (oracle fp64 vs ref)", omel.log_mel_spectrogram(chirp, n), rc, 1.5e-4)
        mel_gold[f"gauss{n}_samples"] = r.reshape(-1)[SAMPLE_IDX[SAMPLE_IDX < r.size]]
        mel_gold[f"chirp{n}_samples"] = rc.reshape(-1)[SAMPLE_IDX[SAMPLE_IDX < rc.size]]
        mel_gold[f"chirp{n}_stats"] = np.array([rc.min(), rc.max(), rc.mean()], dtype=np.float64)
    # batched semantics: global max across the batch (reference audio.py:159), 60 dB level difference
    two = np.stack([pcm[0], pcm[1] * 1e-3]).astype(np.float32)
    rb = ref.log_mel_spectrogram(torch.from_numpy(two), n_mels=80).numpy()
    check("mel80 batched global-max (oracle vs ref)", omel.log_mel_spectrogram(two, 80), rb, 1e-5)
    mel_gold["batch2_global_samples"] = rb.reshape(-1)[SAMPLE_IDX[SAMPLE_IDX < rb.size // 2].tolist()
                                                        + (rb.size // 2 + SAMPLE_IDX[SAMPLE_IDX < rb.size // 2]).tolist()]
    # padding argument + short clip
    rp = ref.log_mel_spectrogram(torch.from_numpy(pcm[0][:16000]), n_mels=80, padding=4800).numpy()
    check("mel80 short clip + padding (oracle vs ref)", omel.log_mel_spectrogram(pcm[0][:16000], 80, padding=4800), rp, 1e-5)
    mel_gold["short_pad_full"] = rp
    np.savez_compressed(os.path.join(GOLD, "mel.npz"), **mel_gold)

    # ------------------------------------------------------------------ 2. network (tiny + gated x-attn)
    print("== network (tiny, gated x-attn, n_langs=1, bert_dim=1024)")
    dims = ModelDimensions(**TINY)
    model = Whisper(dims, 0.0, False, 256, 1, 1024, 1).eval()
    synth.init_synthetic_(model, seed=0)
    sd = om.cast_state_dict_fp32(model.state_dict())
    odims = om.Dims(**TINY)
    mel1 = ref.log_mel_spectrogram(torch.from_numpy(pcm[:1]), n_mels=80)
    feat = synth.synthetic_features(1, n_frames=100, dim=1024, seed=4321)
    toks = torch.tensor([[50258, 50259, 50359, 50363, 1000, 2000, 3000, 4000]])
    with torch.no_grad():
        t0 = time.time()
        xa_ref = model.encoder(mel1)
        print(f"  reference tiny encoder: {time.time() - t0:.2f} s")
        lg_ref = model.decoder(toks, xa_ref, xt_list=[feat])
        xa_or = om.encoder_forward(sd, odims, mel1)
        lg_or = om.decoder_forward(sd, odims, toks, xa_or, xt_list=[feat])
    check("encoder output (oracle vs ref)", xa_or, xa_ref, 2e-5)
    check("decoder logits (oracle vs ref)", lg_or, lg_ref, 2e-4)
    net_gold = {"xa_samples": xa_ref.reshape(-1)[::4001].numpy(), "xa_row0": xa_ref[0, 0].numpy(),
                "xa_row1499": xa_ref[0, 1499].numpy(), "logits_samples": lg_ref.reshape(-1)[::1009].numpy(),
                "logits_last_top": torch.topk(lg_ref[0, -1], 8).indices.numpy(), "tokens": toks.numpy()}
    # the reference's own bf16 drift (basis for the bf16 tolerances)
    with torch.no_grad():
        xa_b = om.encoder_forward(sd, odims, mel1, act_dtype=torch.bfloat16).float()
        lg_b = om.decoder_forward(sd, odims, toks, xa_b.bfloat16(), xt_list=[feat])
    rel = lambda a, b: float((a - b).norm() / b.norm())
    meta["bf16_drift_tiny"] = {"encoder_rel_l2": rel(xa_b, xa_ref), "logits_rel_l2": rel(lg_b, lg_ref)}
    print(f"  reference-semantics bf16 drift: encoder {meta['bf16_drift_tiny']['encoder_rel_l2']:.2e}, "
          f"logits {meta['bf16_drift_tiny']['logits_rel_l2']:.2e}")
    np.savez_compressed(os.path.join(GOLD, "net_tiny_av.npz"), **net_gold)

    # ------------------------------------------------------------------ 3. decoding
    print("== decoding")
    dec_gold = {}
    # 3a. BASELINE config 1: tiny audio-only, fp32, greedy 64 tokens, through the UNMODIFIED reference decode()
    model_a = Whisper(dims, 0.0, False, 256, 0, 768, 0).eval()
    synth.init_synthetic_(model_a, seed=0)
    sd_a = om.cast_state_dict_fp32(model_a.state_dict())
    opt = DecodingOptions(language="en", without_timestamps=True, sample_len=64, fp16=False)
    t0 = time.time()
    res = model_a.decode(mel1, opt)
    print(f"  reference tiny decode() 64 tokens: {time.time() - t0:.2f} s")
    spec = spec_from_task(DecodingTask(model_a, opt), opt)
    ores = odec.decode(sd_a, odims, spec, mel1)
    assert ores[0].tokens == res[0].tokens, "greedy tokens differ (audio-only)"
    check("avg_logprob greedy audio-only", ores[0].avg_logprob, res[0].avg_logprob, 1e-5)
    check("no_speech_prob greedy audio-only", ores[0].no_speech_prob, res[0].no_speech_prob, 1e-6)
    print(f"  [ok] config-1 greedy tokens identical ({len(res[0].tokens)} tokens)")
    dec_gold["greedy_audio_only"] = {"spec": spec_to_json(spec), "tokens": res[0].tokens, "avg_logprob": res[0].avg_logprob,
                                     "no_speech_prob": res[0].no_speech_prob, "text": res[0].text}
    # 3b. greedy with timestamps (ApplyTimestampRules) and default sample_len capped for time
    opt_ts = DecodingOptions(language="en", without_timestamps=False, sample_len=24, fp16=False)
    res_ts = model_a.decode(mel1, opt_ts)
    spec_ts = spec_from_task(DecodingTask(model_a, opt_ts), opt_ts)
    ores_ts = odec.decode(sd_a, odims, spec_ts, mel1)
    assert ores_ts[0].tokens == res_ts[0].tokens, "greedy tokens differ (timestamps)"
    print(f"  [ok] timestamp-rule greedy tokens identical: {res_ts[0].tokens[:8]}...")
    dec_gold["greedy_timestamps"] = {"spec": spec_to_json(spec_ts), "tokens": res_ts[0].tokens,
                                     "avg_logprob": res_ts[0].avg_logprob}
    # 3c. beam search (reference works for one audio only), audio-only
    opt_b = DecodingOptions(language="en", without_timestamps=True, sample_len=16, beam_size=3, fp16=False)
    res_b = model_a.decode(mel1, opt_b)
    spec_b = spec_from_task(DecodingTask(model_a, opt_b), opt_b)
    ores_b = odec.decode(sd_a, odims, spec_b, mel1)
    assert ores_b[0].tokens == res_b[0].tokens, "beam tokens differ"
    check("avg_logprob beam", ores_b[0].avg_logprob, res_b[0].avg_logprob, 1e-5)
    print(f"  [ok] beam-3 tokens identical: {res_b[0].tokens[:8]}...")
    dec_gold["beam3_audio_only"] = {"spec": spec_to_json(spec_b), "tokens": res_b[0].tokens,
                                    "avg_logprob": res_b[0].avg_logprob}
    # 3d. AV greedy: 2 clips, gated x-attn, features 100 x 1024
    mel2 = torch.stack([ref.log_mel_spectrogram(torch.from_numpy(pcm[i]), n_mels=80) for i in range(2)])
    feat2 = synth.synthetic_features(2, n_frames=100, dim=1024, seed=4321)
    opt_av = DecodingOptions(language="en", without_timestamps=True, sample_len=24, fp16=False)
    task_av = DecodingTask(model, opt_av)
    av_ref, nsp_ref = reference_av_greedy(model, task_av, mel2, feat2)
    spec_av = spec_from_task(task_av, opt_av)
    ores_av = odec.decode(sd, odims, spec_av, mel2, feat2)
    for i in range(2):
        assert ores_av[i].tokens == av_ref[i][0], f"AV greedy tokens differ (clip {i})"
        check(f"avg_logprob AV clip {i}", ores_av[i].avg_logprob, av_ref[i][1], 1e-5)
    # the features must matter: audio-only weights on the same mel give different tokens
    print(f"  [ok] AV greedy tokens identical for 2 clips: {av_ref[0][0][:6]}... / {av_ref[1][0][:6]}...")
    dec_gold["greedy_av"] = {"spec": spec_to_json(spec_av), "tokens": [a[0] for a in av_ref],
                             "avg_logprob": [a[1] for a in av_ref], "no_speech_prob": nsp_ref,
                             "feat_frames": 100, "feat_dim": 1024}
    # 3e. long-form transcribe(): 75 s of synthetic audio, 3 windows, temperature 0, thresholds
    # off (no random fallback), 16 sampled tokens per window, conditioning on the previous text
    from ref_whisper.transcribe import transcribe as ref_transcribe
    long_pcm = synth.synthetic_pcm(1, n_samples=75 * 16000, seed=99)[0].numpy()
    tr = ref_transcribe(model_a, long_pcm, temperature=0.0, compression_ratio_threshold=None, logprob_threshold=None,
                        no_speech_threshold=None, language="en", fp16=False, sample_len=16, verbose=None)
    print(f"  reference transcribe(): {len(tr['segments'])} segments, seeks {[s['seek'] for s in tr['segments']]}")
    dec_gold["transcribe_long"] = {
        "seconds": 75, "seed": 99, "text": tr["text"], "language": tr["language"],
        "segments": [{k: (s[k] if k != "tokens" else list(map(int, s[k]))) for k in
                      ("id", "seek", "start", "end", "text", "tokens", "temperature", "avg_logprob", "no_speech_prob")}
                     for s in tr["segments"]]}
    # 3f. detect_language (decoding.py:18-77) on the audio-only model, 2 clips; and decode() with language=None
    ltoks, lprobs = model_a.detect_language(mel2)
    tk_a = DecodingTask(model_a, opt).tokenizer
    lang_ids = list(tk_a.all_language_tokens)
    xa2_or = om.encoder_forward(sd_a, odims, mel2)
    o_best, o_probs = odec.detect_language(sd_a, odims, xa2_or, tk_a.sot, lang_ids)
    assert o_best == ltoks.tolist(), "detected language tokens differ"
    for i in range(2):
        check(f"language probs clip {i}", o_probs[i].numpy(), [lprobs[i][c] for c in tk_a.all_language_codes], 1e-6)
    opt_auto = DecodingOptions(language=None, without_timestamps=True, sample_len=12, fp16=False)
    res_auto = model_a.decode(mel2, opt_auto)
    print(f"  [ok] detect_language: tokens {ltoks.tolist()} -> {[r.language for r in res_auto]}")
    dec_gold["detect_language"] = {
        "language_tokens": ltoks.tolist(), "languages": [r.language for r in res_auto],
        "top_prob": [max(p.values()) for p in lprobs],
        "probs_en": [p["en"] for p in lprobs],
        "auto_tokens": [r.tokens for r in res_auto], "auto_avg_logprob": [r.avg_logprob for r in res_auto]}
    # oracle decode with the detected language spliced into the prompt, clip by clip
    for i in range(2):
        init = list(spec.initial_tokens)
        init[spec.sot_index + 1] = int(ltoks[i])
        sp = odec.DecodeSpec(**{**spec.__dict__, "initial_tokens": tuple(init), "sample_len": 12})
        got = odec.decode(sd_a, odims, sp, mel2[i:i + 1])
        assert got[0].tokens == res_auto[i].tokens, f"language=None decode differs (clip {i})"
    print("  [ok] decode(language=None) tokens identical to the oracle with the detected language token")
    # 3g. multi-feature gated x-attention (num_langs = 3, the fork's real calling mode: trilingual.py:256,304):
    # teacher-forced logits and a greedy loop with xt_list of 3 tensors of different lengths / widths
    model_m = Whisper(dims, 0.0, False, 256, 1, 1024, 3).eval()
    synth.init_synthetic_(model_m, seed=0)
    sd_m = om.cast_state_dict_fp32(model_m.state_dict())
    feats3 = [synth.synthetic_features(2, n_frames=n, dim=w, seed=4321 + j)
              for j, (n, w) in enumerate(((100, 1024), (37, 1024), (64, 384)))]
    toks2 = toks.repeat(2, 1)
    with torch.no_grad():
        xa_m = model_m.encoder(mel2)
        lg_m = model_m.decoder(toks2, xa_m, xt_list=feats3)
        lg_m2 = model_m.decoder(toks2, xa_m, xt_list=feats3[:2])   # fewer tensors than num_langs is allowed
        lg_mo = om.decoder_forward(sd_m, odims, toks2, om.encoder_forward(sd_m, odims, mel2), xt_list=feats3)
    check("multi-feature decoder logits (oracle vs ref)", lg_mo, lg_m, 2e-4)
    task_m = DecodingTask(model_m, opt_av)
    task_m.decoder.reset()
    tokens_m = torch.tensor([task_m.initial_tokens]).repeat(2, 1)
    sum_lp_m = torch.zeros(2)
    with torch.no_grad():
        for i in range(12):
            lgs = model_m.decoder(tokens_m, xa_m, xt_list=feats3)[:, -1]
            for f in task_m.logit_filters:
                f.apply(lgs, tokens_m)
            tokens_m, done = task_m.decoder.update(tokens_m, lgs, sum_lp_m)
    spec_m = odec.DecodeSpec(**{**spec_av.__dict__, "sample_len": 12})
    ores_m = odec.decode(sd_m, odims, spec_m, mel2, feats3)
    ref_m = [row[task_m.sample_begin:] for row in tokens_m.tolist()]
    assert [r.tokens for r in ores_m] == ref_m, "multi-feature greedy tokens differ"
    print(f"  [ok] multi-feature (3 tensors) greedy tokens identical: {ref_m[0][:6]}...")
    net_multi = {"logits3_samples": lg_m.reshape(-1)[::1009].numpy(), "logits2_samples": lg_m2.reshape(-1)[::1009].numpy(),
                 "tokens": toks2.numpy()}
    np.savez_compressed(os.path.join(GOLD, "net_tiny_multi.npz"), **net_multi)
    dec_gold["greedy_multi3"] = {"spec": spec_to_json(spec_m), "tokens": ref_m,
                                 "feat_shapes": [[100, 1024], [37, 1024], [64, 384]], "feat_seeds": [4321, 4322, 4323]}
    with open(os.path.join(GOLD, "decode_tiny.json"), "w") as fh:
        json.dump({"meta": meta, "dims": TINY, "cases": dec_gold}, fh, indent=1)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
