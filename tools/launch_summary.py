#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last step."""
import collections
import csv
import re
import sys


def main(path, steps=4, top=22):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    rows = [(row[ki], float(row[vi].replace(",", ""))) for row in r if len(row) > vi]
    n = len(rows) // steps
    marks = [i for i, (k, _) in enumerate(rows) if "tc_embed_kernel" in k or "ts_embed_kernel" in k]
    if len(marks) >= 2:          # one patch-embedding launch per step: its spacing is the step length
        n = marks[-1] - marks[-2]
    last = rows[-n:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, ns in last:
        k = re.sub(r"\(.*", "", k)[:64]
        agg[k][0] += 1
        agg[k][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"launches in last step: {n}; serialized device time {tot / 1e6:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}% x{v[0]:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4, int(sys.argv[3]) if len(sys.argv) > 3 else 22)
