#!/usr/bin/env python
"""Sum dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum over the launches of ONE step that match a kernel regex,
from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of `bench.py --steps 1
--warmup W --only-resident` (every step launches the same kernels: the last 1/(W+1) of the matching launches is one step).
Prints a tools/ncu_summary.py-style block that bench.py's profile_traffic() parses.
Usage: ncu_dram_total.py log.csv 'regex' steps_in_log [label]"""
import collections
import csv
import re
import sys


def main(path, pattern, steps, label):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, mi, ui, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
    allk = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        d = allk.setdefault(row[ii], {"k": row[ki]})
        d[row[mi]] = float(row[vi].replace(",", "")) * unit.get(row[ui], 1.0)
    ids = list(allk)
    # one patch-embedding launch per step: the launches after the last one are the last step (fallback: 1/steps of the log)
    marks = [i for i, k in enumerate(ids) if "tc_embed_kernel" in allk[k]["k"] or "ts_embed_kernel" in allk[k]["k"]]
    if len(marks) >= 2:
        step_ids = ids[marks[-2]:marks[-1]]        # a complete step: [embed of step k, embed of step k+1)
    else:
        step_ids = ids[-(len(ids) // steps):]
    last = [allk[i] for i in step_ids if re.search(pattern, allk[i]["k"])]
    n = len(last)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in last)
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in last)
    ns = sum(d.get("gpu__time_duration.sum", 0.0) for d in last)
    print(f"kernel: {label} ({n} launches of one step matching /{pattern}/; per-launch ncu replay, cold L2)")
    print(f"  dram__bytes_read.sum   {rd / 1e6:.6f} Mbyte")
    print(f"  dram__bytes_write.sum   {wr / 1e6:.6f} Mbyte")
    print(f"  gpu__time_duration.sum   {ns / 1e3:.3f} us")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in last:
        k = re.sub(r"\(.*", "", d["k"])[:60]
        a = agg[k]
        a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
    print("# per kernel: launches, us, MB read, MB written")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"#   {k:60s} {a[0]:3d} {a[1] / 1e3:9.1f} {a[2] / 1e6:9.1f} {a[3] / 1e6:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "total")
