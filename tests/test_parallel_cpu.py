"""N > 1 host logic on CPU: two gloo ranks exercise the sharding + final result gather (the only collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "whisper-flamingo_b200"))
    from whisper import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n_items, rank, world)
        # rank-dependent ragged rows: row i holds i+1 copies of the value 100+i
        width = max(1, hi)
        local = torch.full((hi - lo, width), -1, dtype=torch.int32)
        for r, i in enumerate(range(lo, hi)):
            local[r, : i + 1] = 100 + i
        full = parallel.gather_token_matrix(local, -1)
        floats = parallel.gather_floats([float(i) * 0.5 for i in range(lo, hi)], "cpu")
        out_q.put((rank, full.tolist(), floats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 2, 1])
def test_two_rank_gloo_gather_matches_single_process(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_rows = [[100 + i] * (i + 1) for i in range(n_items)]
    for rank, full, floats in results:
        assert len(full) == n_items
        for row, want in zip(full, want_rows):
            assert row[: len(want)] == want and all(v == -1 for v in row[len(want):])
        assert floats == [i * 0.5 for i in range(n_items)]  # rank order == clip order on every rank


def test_shard_range_covers_everything_once():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "whisper-flamingo_b200"))
    from whisper.parallel import shard_range
    for n in (0, 1, 7, 8, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_transcribe_batch_row_buckets():
    """Host logic of the long-form driver: batch sizes are padded to few distinct values so that decode sessions and
    CUDA graphs are reused while the batch shrinks (whisper/transcribe.py:_bucket_rows)."""
    from whisper.transcribe import _bucket_rows
    assert [_bucket_rows(n, 16, 128) for n in (1, 2, 3, 5, 8, 9, 15, 16, 17, 33, 64, 113, 128)] == \
        [1, 2, 4, 8, 8, 16, 16, 16, 32, 48, 64, 128, 128]
    assert _bucket_rows(120, 16, 120) == 120          # the bucket would exceed max_batch: no padding
    assert _bucket_rows(7, 1, 128) == 7 and _bucket_rows(0, 16, 128) == 0
    for n in range(1, 129):
        b = _bucket_rows(n, 16, 128)
        assert n <= b <= max(2 * n, n + 15)
