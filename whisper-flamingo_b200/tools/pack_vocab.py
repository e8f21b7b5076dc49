"""Repack an OpenAI-Whisper BPE vocabulary (``*.tiktoken``: one ``base64(token) rank``
pair per line) into the compact binary the drop-in tokenizer loads.

Output format (gzip):  u32 n_tokens, then for rank 0..n-1:  u16 byte_length, raw bytes.
The vocabularies themselves are OpenAI Whisper data assets (MIT licence); they are data,
not code, and cannot be regenerated, so they are carried in repacked form.

usage: python pack_vocab.py <multilingual.tiktoken> <out.vocab.gz>
"""
import base64
import gzip
import struct
import sys


def main(src: str, dst: str) -> None:
    ranks = {}
    with open(src) as fh:
        for line in fh:
            if line.strip():
                tok, rank = line.split()
                ranks[int(rank)] = base64.b64decode(tok)
    n = len(ranks)
    assert sorted(ranks) == list(range(n)), "ranks must be dense"
    blob = bytearray(struct.pack("<I", n))
    for r in range(n):
        b = ranks[r]
        blob += struct.pack("<H", len(b)) + b
    with gzip.GzipFile(dst, "wb", mtime=0) as fh:
        fh.write(bytes(blob))
    print(f"{src} -> {dst}: {n} tokens, {len(blob)} bytes raw")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
