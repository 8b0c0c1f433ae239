#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libstep_b200.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG = TMA bulk / tensor copies, UTCBAR = tcgen05.commit,
UTCATOMSWS = tcgen05.alloc, SYNCS = mbarrier ops, LDG.E.256 = 256-bit global loads, HMMA = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PATS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDG.E.*256", "HMMA", "FFMA2", "MUFU.EX2"]


def main():
    lib = os.path.join(ROOT, "step_b200", "libstep_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("stepk::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PATS:
            if re.search(r"\b" + p + r"\b" if p.isalnum() else p, line):
                counts[cur][p] += 1
    hdr = "kernel".ljust(52) + "".join(p.replace(".E.*", "").rjust(11) for p in PATS)
    print(hdr)
    tot = collections.Counter()
    for k, c in counts.items():
        if any(c[p] for p in PATS[:8]):
            print(k[:51].ljust(52) + "".join(str(c[p]).rjust(11) for p in PATS))
        tot.update(c)
    print("TOTAL (all %d kernels)" % len(counts) + " " * 28 + "".join(str(tot[p]).rjust(11) for p in PATS)[6:])


if __name__ == "__main__":
    main()
