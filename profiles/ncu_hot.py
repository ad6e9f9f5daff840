#!/usr/bin/env python
"""profiles/ncu_hot.py <report.ncu-rep> [trials] — SASS-level hot spots from the source page of an
ncu capture: executed warp-instructions per opcode and the instruction ranges that execute most."""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    trials = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samples = Counter(), Counter()
    tot = 0
    body = []
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        n = int(r[iex] or 0)
        sm = int(r[ismp] or 0)
        src = r[isrc].strip()
        op = src.split()[0]
        if op.startswith("@"):
            op = src.split()[1]
        ops[op.rstrip(";")] += n
        samples[op.rstrip(";")] += sm
        tot += n
        body.append((n, sm, src))
    print(f"total warp-instructions {tot}" + (f"  = {tot / trials:.1f} per trial" if trials else ""))
    for op, n in ops.most_common(28):
        extra = f" {n / trials:8.1f}/trial" if trials else ""
        print(f"  {op:24s} {n:14d} {100 * n / tot:5.1f}%{extra}  samples {samples[op]}")
    # execution-count profile: how many static instructions run how often (per trial)
    if trials:
        buckets = Counter()
        for n, sm, src in body:
            buckets[round(n / trials, 1)] += 1
        print("static instructions by executions/trial:", sorted(buckets.items(), key=lambda kv: -kv[0] * kv[1])[:12])


if __name__ == "__main__":
    main()
