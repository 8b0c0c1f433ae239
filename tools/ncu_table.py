#!/usr/bin/env python
"""Aggregate a tools/ncu_summary.py text (many launches) into one line per kernel: launches, total time, DRAM bytes,
issue / FMA / tensor pipe utilisation (averages over the launches)."""
import collections
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path):
    blocks = open(path).read().split("kernel:")[1:]
    agg = collections.OrderedDict()
    for b in blocks:
        lines = b.strip().split("\n")
        name = re.sub(r"\(.*", "", lines[0]).strip()[:56]
        v = {}
        for l in lines[1:]:
            p = l.split()
            if len(p) >= 2:
                try:
                    v[p[0]] = float(p[1].replace(",", "")) * (UNIT.get(p[2], 1.0) if len(p) > 2 else 1.0)
                except ValueError:
                    pass
        a = agg.setdefault(name, collections.defaultdict(float))
        a["n"] += 1
        a["us"] += v.get("gpu__time_duration.sum", 0.0)
        a["rd"] += v.get("dram__bytes_read.sum", 0.0)
        a["wr"] += v.get("dram__bytes_write.sum", 0.0)
        for k, key in (("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       ("fma", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                       ("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                       ("l1", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                       ("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):
            a[k] += v.get(key, 0.0)
        a["regs"] = v.get("launch__registers_per_thread", 0.0)
    print(f"{'kernel':56s} {'n':>3s} {'us tot':>8s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'issue%':>6s} {'fma%':>6s} {'tens%':>6s} {'l1%':>5s} regs")
    tot = collections.defaultdict(float)
    for k, a in agg.items():
        n = a["n"]
        print(f"{k:56s} {int(n):3d} {a['us']:8.1f} {a['rd']:8.1f} {a['wr']:8.1f} {a['dram'] / n:6.1f} {a['issue'] / n:6.1f} "
              f"{a['fma'] / n:6.1f} {a['tensor'] / n:6.1f} {a['l1'] / n:5.1f} {int(a['regs'])}")
        for x in ("us", "rd", "wr"):
            tot[x] += a[x]
    print(f"{'TOTAL':56s}     {tot['us']:8.1f} {tot['rd']:8.1f} {tot['wr']:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
