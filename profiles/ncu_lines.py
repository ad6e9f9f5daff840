#!/usr/bin/env python
"""profiles/ncu_lines.py <report.ncu-rep> <cubin.asm from `nvdisasm -g -c`> <kernel-substring> [trials]
Executed warp-instructions and stall samples per CUDA SOURCE LINE: joins ncu's per-SASS-instruction
counters (source page, in program order) with nvdisasm's line markers for the same function."""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_lines(asm, kern):
    """yield (opcode text, 'file:line') for each instruction of the kernel, in order"""
    out, on, cur = [], False, "?"
    for ln in open(asm):
        if ln.startswith(".text.") or ln.lstrip().startswith(".section"):
            on = (kern in ln) and ".text." in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
            # inlined-from chains: keep the innermost marker
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), cur))
    return out


def main():
    rep, asm, kern = sys.argv[1:4]
    trials = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    prof = [(r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
    sl = sass_lines(asm, kern)
    if len(sl) != len(prof):
        print(f"warning: {len(sl)} disassembled vs {len(prof)} profiled instructions; joining by order up to the shorter")
    ex, smp, fp64 = defaultdict(int), defaultdict(int), defaultdict(int)
    for (op, line), (src, n, s) in zip(sl, prof):
        ex[line] += n
        smp[line] += s
        o = src.split()[1] if src.startswith("@") else src.split()[0]
        if o.startswith(("DFMA", "DMUL", "DADD", "DSETP")):
            fp64[line] += n
    tot, tots = sum(ex.values()), sum(smp.values())
    print(f"total {tot / trials:.1f} warp-instr/trial, {sum(fp64.values()) / trials:.1f} FP64-pipe/trial")
    for line, n in sorted(ex.items(), key=lambda kv: -kv[1])[:400]:
        print(f"  {line:28s} {n / trials:8.1f} instr/trial  fp64 {fp64[line] / trials:6.1f}  stall-samples {100 * smp[line] / max(1, tots):5.1f}%")


if __name__ == "__main__":
    main()
