#!/usr/bin/env python
"""profiles/sanitize_case.py — a small run that touches every kernel (STRICT and FAST sweeps, fed and Philox,
block sweep, all-particle with half shell / full shell / cluster, evaluate, gather) for compute-sanitizer:
   compute-sanitizer --tool memcheck  python profiles/sanitize_case.py
   compute-sanitizer --tool racecheck python profiles/sanitize_case.py      (one tool per gpurun call)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
smcb = importlib.import_module("montecarlo-surfacer_b200")

rng = np.random.default_rng(0)
W = smcb.REFERENCE_WALL_M3


def droplet(N, Lz):
    n = int(np.ceil((N / 4) ** 0.5))
    g = np.array([(i, j, k) for k in range(4) for i in range(n) for j in range(n)], dtype=float)[:N]
    g[:, :2] = (g[:, :2] - n / 2) * 1.12
    g[:, 2] = -Lz / 2 + 0.95 + g[:, 2] * 1.12
    return (g + (rng.random(g.shape) - .5) * 0.1).reshape(-1)


for N, Lz in ((108, 200.0), (40, 200.0)):
    with smcb.Engine(3, N, 3) as eng:
        eng.set_params(smcb.default_params(L=33.0, Lz=Lz, T=1.1, A=0.05), W)
        eng.set_positions(np.stack([droplet(N, Lz) for _ in range(3)]))
        eng.set_rng(1, 0, 0)
        eng.sweep(3, smcb.FAST); eng.sweep(2, smcb.STRICT)
        S = 2
        eng.sweep_fed(rng.standard_normal((S, 3, 3 * N)) * 0.3, rng.integers(0, 2 ** 31 - 1, (S, 3)), rng.random((S, 3, N)), mode=smcb.FAST)
        eng.sweep_fed(rng.standard_normal((S, 3, 3 * N)) * 0.3, rng.integers(0, 2 ** 31 - 1, (S, 3)), rng.random((S, 3, N)), mode=smcb.STRICT)
        eng.set_params(smcb.default_params(L=33.0, Lz=Lz, T=1.1, A=1e-4), W)
        eng.step_allparticle(4, smcb.FAST); eng.step_allparticle(2, smcb.STRICT)
        eng.evaluate(smcb.FAST); eng.evaluate(smcb.STRICT); eng.gather(); eng.obs_get()
os.environ["SMCB_CLUSTER"] = "2"
with smcb.Engine(2, 600, 3) as eng:                       # block sweep + clustered / full-shell all-particle
    eng.set_params(smcb.default_params(L=33.0, Lz=240.0, T=1.1, A=0.02), W)
    eng.set_positions(np.stack([droplet(600, 240.0) for _ in range(2)]))
    eng.set_rng(2, 0, 0)
    eng.sweep(1, smcb.FAST)
    eng.set_params(smcb.default_params(L=33.0, Lz=240.0, T=1.1, A=1e-5), W)
    eng.step_allparticle(2, smcb.FAST)
    eng.evaluate(smcb.FAST); eng.gather()
print("sanitize case done")
