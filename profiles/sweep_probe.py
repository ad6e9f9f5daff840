#!/usr/bin/env python
"""profiles/sweep_probe.py — the FAST sweep kernel alone on the three states of the headline batch
(8192 chains x N=256, main.c geometry): the start lattice (as bench.py times it), the thermalised gas
(2000 sweeps with 2A first) and the condensed droplet on the wall (A = 0.02).  Kernel time only (CUDA
events on the engine's stream), one JSON line per state.  SMCB_SWEEP_KERNEL=cached selects the
first-generation kernel for A/B runs.  Run on the GPU box.

  python profiles/sweep_probe.py [--chains 8192] [--states lattice,thermal,droplet] [--reps 3]
"""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
smcb = importlib.import_module("montecarlo-surfacer_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=8192)
ap.add_argument("--states", default="lattice,thermal,droplet")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sweeps", type=int, default=40)
ap.add_argument("--N", type=int, default=256)
args = ap.parse_args()
N, C, S = args.N, args.chains, args.sweeps
L, Lz, T = 33.0, 240.0, 1.1


def lattice():
    nxy, nz = (4, 4) if N == 256 else (3, 3)
    a = L / nxy
    cells = np.array([(i, j, k) for i in range(nxy) for j in range(nxy) for k in range(nz)], dtype=float)
    basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    X = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + a / 4
    Pz = Lz - Lz / 20.0
    X[:, :2] -= L * np.rint(X[:, :2] / L)
    X[:, 2] -= Pz * np.rint(X[:, 2] / Pz)
    return X.reshape(-1)


def droplet():
    nzl = 4
    nxy = int(np.ceil(np.sqrt(N / nzl)))
    g = np.array([(i, j, k) for k in range(nzl) for i in range(nxy) for j in range(nxy)], dtype=float)[:N]
    g[:, 0] = (g[:, 0] - nxy / 2) * 1.12
    g[:, 1] = (g[:, 1] - nxy / 2) * 1.12
    g[:, 2] = -Lz / 2 + 0.95 + g[:, 2] * 1.12
    rs = np.random.default_rng(7)
    g += (rs.random(g.shape) * 2 - 1) * 0.05
    return g[rs.permutation(N)].reshape(-1)


variant = os.environ.get("SMCB_SWEEP_KERNEL", "auto")
for state in args.states.split(","):
    A = 0.02 if state == "droplet" else T
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), smcb.REFERENCE_WALL_M3)
        eng.broadcast_positions(droplet() if state == "droplet" else lattice())
        eng.set_rng(12345, 0, 0)
        if state == "thermal":
            eng.set_step_scale(2.0)
            eng.sweep(2000, smcb.FAST)
            eng.set_step_scale(1.0)
        for _ in range(3):
            eng.sweep(S, smcb.FAST)
        eng.reset_counters()
        ms = []
        for _ in range(args.reps):
            eng.sweep(S, smcb.FAST)
            ms.append(eng.last_kernel_ms()[0])
        tot, cut = eng.last_pair_counts()
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        Erec = ev["U_lj"] + ev["U_wall"]
        drift = float(np.max(np.abs(E - Erec) / np.maximum(1.0, np.abs(Erec))))
        m = float(np.mean(ms))
        print(json.dumps({"kernel": variant, "state": state, "chains": C, "N": N, "ms_per_launch": m, "ms_all": ms,
                          "pair_int_per_s": C * S * 2.0 * N * (N - 1) / (m * 1e-3), "sweeps_per_s": C * S / (m * 1e-3),
                          "in_cutoff_frac": cut / max(1, tot), "acceptance": float(na.sum()) / max(1, int(nt.sum())),
                          "E_mean": float(E.mean()), "running_E_vs_recomputed": drift}), flush=True)
