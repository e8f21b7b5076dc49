"""Small host-side helpers on the decode path (reference whisper/utils.py:24-26, 45-47)."""
import zlib


def exact_div(x: int, y: int) -> int:
    assert x % y == 0
    return x // y


def compression_ratio(text: str) -> float:
    raw = text.encode("utf-8")
    return len(raw) / len(zlib.compress(raw))
