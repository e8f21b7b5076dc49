"""Data-parallel helpers: clips are independent, so every rank (one process per GPU) holds a full replica and
decodes its own contiguous block of clips; the ONLY collective is the final gather of the results (token ids
int32 ``[B, T]`` + three floats per clip, a few KB) - NCCL over NVLink on the GPU box, gloo in the CPU tests.
There is no counterpart in the reference (its inference is single-process)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_token_matrix(tokens: torch.Tensor, pad_value: int, group=None) -> torch.Tensor:
    """all_gather of ragged per-rank int32 matrices [b_r, t_r] -> [sum b_r, max t_r] in rank order
    (rows padded with pad_value).  Works with any backend (tensor stays on its device)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tokens
    world = dist.get_world_size(group)
    shape = torch.tensor(list(tokens.shape), dtype=torch.int64, device=tokens.device)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    b_max = int(max(s[0] for s in shapes))
    t_max = int(max(s[1] for s in shapes))
    padded = torch.full((b_max, t_max), pad_value, dtype=tokens.dtype, device=tokens.device)
    padded[: tokens.shape[0], : tokens.shape[1]] = tokens
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: int(s[0])] for o, s in zip(out, shapes)], dim=0)


def gather_floats(values: Sequence[float], device, group=None) -> List[float]:
    """Concatenate one float list per rank in rank order (lists may differ in length)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device).view(-1, 1)
    g = gather_token_matrix(t, float("nan"), group)
    return g.view(-1).tolist()


def decode_sharded(model, mel: torch.Tensor, options, x_v: Optional[torch.Tensor] = None, group=None):
    """whisper.decode over the ranks of `group`: every rank passes the SAME full batch, decodes its own block and
    receives (tokens [B, T] padded with EOT, avg_logprob [B], no_speech_prob [B]) for the whole batch."""
    from .decoding import DecodingTask, decode
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(mel.shape[0], rank, world)
    eot = DecodingTask(model, options).tokenizer.eot
    if hi > lo:
        res = decode(model, mel[lo:hi], options, x_v=None if x_v is None else x_v[lo:hi])
        t_max = max(1, max(len(r.tokens) for r in res))
        toks = torch.full((hi - lo, t_max), eot, dtype=torch.int32, device=mel.device)
        for i, r in enumerate(res):
            toks[i, : len(r.tokens)] = torch.tensor(r.tokens, dtype=torch.int32)
        lps, nsp = [r.avg_logprob for r in res], [r.no_speech_prob for r in res]
    else:
        toks = torch.full((0, 1), eot, dtype=torch.int32, device=mel.device)
        lps, nsp = [], []
    return (gather_token_matrix(toks, eot, group), gather_floats(lps, mel.device, group),
            gather_floats(nsp, mel.device, group))
