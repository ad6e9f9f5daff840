#!/usr/bin/env python
"""profiles/launch_list.py <ncu --csv launch log> — per-launch device times (kernel, grid, block, ns)
and each kernel's share of the total, from the `--metrics gpu__time_duration.sum` pass."""
import csv
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, ig, ib, iv = (hdr.index(x) for x in ("Kernel Name", "Grid Size", "Block Size", "Metric Value"))
    tot, per = 0.0, defaultdict(lambda: [0, 0.0])
    print("id,kernel,grid,block,ns")
    for n, r in enumerate(rows[1:]):
        name = r[ik].split("(")[0].replace("void ", "")
        ns = float(r[iv].replace(",", ""))
        print(f'{n},"{name}","{r[ig]}","{r[ib]}",{ns:.0f}')
        per[name][0] += 1
        per[name][1] += ns
        tot += ns
    print("# share of total device time (cold-cache, serialised launches: compare shares, not absolutes)")
    for name, (cnt, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"# {100 * ns / tot:6.2f}%  {ns / 1e6:10.3f} ms  x{cnt:<3d} {name}")


if __name__ == "__main__":
    main()
