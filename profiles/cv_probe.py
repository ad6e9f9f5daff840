#!/usr/bin/env python
"""profiles/cv_probe.py - energy fluctuation of the FAST sweep as sMC measures it (cv = var(E[0..maxsteps]) / T^2 over one
chain's series, SMC.c:250), on many chains: N = 108, main.c's box, T = A = 1.1, 400 sweeps from the start lattice
(eq = 0) or after 100 sweeps at 2A (eq = 100).  The CPU side of the comparison (the oracle's sweep on 1536 chains
per arm, reference stream and independent Gaussians alike): cv = 0.7245 +- 0.0049 (eq = 0), 0.7074 +- 0.014 (eq = 100).
Run on the GPU box."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
smcb = importlib.import_module("montecarlo-surfacer_b200")
from oracle_bindings import GOLDEN_W_M3, Oracle
N, M, T, A, L, LZ, C, S = 108, 3, 1.1, 1.1, 33.0, 200.0, 8192, 400
R0, _ = Oracle().initialize_box(L, LZ, N)
for kernel in ("auto", "cached"):
    os.environ["SMCB_SWEEP_KERNEL"] = kernel
    for eq in (0, 100):
        with smcb.Engine(C, N, M) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=LZ, T=T, A=A), GOLDEN_W_M3)
            eng.broadcast_positions(R0)
            eng.set_rng(31337 + eq, 0, 0)
            eng.refresh_energy(smcb.FAST)
            if eq:
                eng.set_step_scale(2.0); eng.sweep(eq, smcb.FAST); eng.set_step_scale(1.0)
            E0 = eng.chain_state()[0].copy()
            tr = eng.sweep_traced(S, smcb.FAST)
            Es = np.concatenate([E0[None, :], np.asarray(tr[0]).reshape(S, C)], axis=0)
            cv = Es.var(axis=0) / T**2
            print(json.dumps({"kernel": kernel, "eq": eq, "chains": C, "cv": cv.mean(), "cv_se": cv.std(ddof=1) / np.sqrt(C),
                              "E_mean": Es.mean(), "E_se": Es.mean(axis=0).std(ddof=1) / np.sqrt(C)}), flush=True)
