"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python tools/ncu_launches.py launches.csv [first_decode_launch_index]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        us = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3) * v
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:wf::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
        short = (m.group(1) + (m.group(2) or "")) if m else name[:50]
        grid = row.get("Grid Size", "")
        out.append((short, us, grid))
    return out


def table(rows, title):
    agg = collections.OrderedDict()
    for name, us, _ in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"## {title}: {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time")
    print(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {c:8d} {t / 1e3:10.3f} {t / c:9.2f} {100 * t / tot:6.1f}%")
    return tot


if __name__ == "__main__":
    rows = load(sys.argv[1])
    table(rows, "all captured launches")
    # the decode loop starts at the first embed_kernel
    idx = [i for i, r in enumerate(rows) if r[0].startswith("embed_kernel")]
    if idx:
        table(rows[: idx[0]], "log-mel + encoder + per-clip K/V precompute")
        if len(idx) > 2:
            step = rows[idx[1]: idx[2]]
            table(step, "one decode step (second step captured)")
