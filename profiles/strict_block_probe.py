import importlib, os, sys, time, json
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
smcb = importlib.import_module("montecarlo-surfacer_b200")
from oracle_bindings import GOLDEN_W_M3, Oracle
N, L, Lz = 4096, 33.0, 240.0
X = Oracle().fcc_lattice(L, Lz, 16, 16, 4)
for C in (1, 148):
    for which in ("serial", "spec"):
        os.environ["SMCB_BLOCK_SWEEP"] = which
        with smcb.Engine(C, N, 3) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=0.05), GOLDEN_W_M3)
            eng.broadcast_positions(X)
            eng.set_rng(5, 0, 0)
            eng.refresh_energy(smcb.STRICT)
            eng.sweep(1, smcb.STRICT)
            eng.sweep(2, smcb.STRICT)
            ms = eng.last_kernel_ms()[0]
            E, na, nt = eng.chain_state()
            print(json.dumps({"strict_block": which, "chains": C, "ms_per_sweep": ms / 2, "sweeps_per_s_per_chain": 2e3 / ms, "acceptance": float(na.sum() / nt.sum()), "E0": float(E[0])}), flush=True)
