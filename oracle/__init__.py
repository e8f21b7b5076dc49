"""CPU oracle for the Whisper-Flamingo inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product path
(``whisper-flamingo_b200/whisper``) never imports this package and fails
loudly when its CUDA library is missing.

What is restated here (reference = jerryyang1231/whisper-flamingo, paths
relative to the reference root):

* ``oracle.mel``     - ``whisper/audio.py:66-161``  (pad_or_trim, mel_filters,
                       log_mel_spectrogram)
* ``oracle.model``   - ``whisper/model.py:30-340``  (LayerNorm/Linear/Conv1d
                       dtype policy, MultiHeadAttention, GatedXAttnSubBlock,
                       ResidualAttentionBlock, AudioEncoder, TextDecoder)
* ``oracle.decode``  - ``whisper/decoding.py:276-509, 688-798`` (greedy / beam
                       token decoders, logit filters, the no-KV-cache main loop,
                       result assembly), with the one argument the reference
                       forgets (``xt_list``) supplied - SURVEY.md Appendix B.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
F12), so the oracle is pinned against the *live* reference imported in the
authoring container by ``oracle/pin_reference.py``; that script also writes
the fixtures under ``tests/golden/`` which travel to the GPU box (the
reference itself does not).  Third-party arithmetic (PyTorch ATen, tiktoken)
is not under /root/reference; parity is therefore defined operationally as
agreement with the reference run on torch 2.11.0 CPU in this image.
"""
