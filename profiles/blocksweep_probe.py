#!/usr/bin/env python
"""profiles/blocksweep_probe.py — where the block-per-chain sweep (N = 4096) spends its time: real cutoff vs a cutoff
so small that no pair is ever inside it (screen + reductions + barriers only).  Run on the GPU box."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
smcb = importlib.import_module("montecarlo-surfacer_b200")
from oracle_bindings import Oracle

N, C, S = 4096, 148, 2
R0 = Oracle().fcc_lattice(33.0, 240.0, 16, 16, 4)
for name, rc2, wall in (("real cutoff rc=3, wall", 9.0, smcb.WALL), ("no partners rc=0.1, wall", 0.01, smcb.WALL), ("no partners, no wall", 0.01, 0)):
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=33.0, Lz=240.0, T=1.1, A=1.1, rc2=rc2, flags=wall), smcb.REFERENCE_WALL_M3)
        eng.broadcast_positions(R0)
        eng.set_rng(12345, 0, 0)
        eng.sweep(S, smcb.FAST)
        ms = []
        for _ in range(3):
            eng.sweep(S, smcb.FAST)
            ms.append(eng.last_kernel_ms()[0])
        E, na, nt = eng.chain_state()
        print(f"{name:26s} {np.mean(ms):7.2f} ms per {S} sweeps  acceptance {na.sum() / nt.sum():.3f}  "
              f"{np.mean(ms) * 1e-3 * 1.965e9 / (S * N):6.0f} cycles per trial")
