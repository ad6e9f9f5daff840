#!/usr/bin/env python
"""profiles/fastpath_probe.py — how the sweep kernel's time splits between its paths: the same 8192 x N=256
batch with (a) the real cutoff and (b) a cutoff so small that no trial ever finds a partner (every trial takes
the speculative fast path).  Run on the GPU box."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
smcb = importlib.import_module("montecarlo-surfacer_b200")
from oracle_bindings import GOLDEN_W_M3, Oracle

N, C, S = 256, 8192, 40
R0, _ = Oracle().initialize_box(33.0, 240.0, N)
for name, rc2 in (("real cutoff rc=3", 9.0), ("no partners rc=0.1", 0.01)):
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=33.0, Lz=240.0, T=1.1, A=1.1, rc2=rc2), GOLDEN_W_M3)
        eng.broadcast_positions(R0)
        eng.set_rng(12345, 0, 0)
        for _ in range(3):
            eng.sweep(S, smcb.FAST)
        ms = []
        for _ in range(3):
            eng.sweep(S, smcb.FAST)
            ms.append(eng.last_kernel_ms()[0])
        tot, cut = eng.last_pair_counts()
        E, na, nt = eng.chain_state()
        cyc = np.mean(ms) * 1e-3 * 1.965e9 / (C * S * N / (148 * 4))
        print(f"{name:22s} {np.mean(ms):7.2f} ms per {S} sweeps  in-cutoff frac {cut / tot:.5f}  acceptance {na.sum() / nt.sum():.3f}  "
              f"{cyc:6.0f} SM-sub-partition cycles per trial")
