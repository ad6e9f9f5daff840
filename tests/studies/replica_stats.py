#!/usr/bin/env python
"""tests/studies/replica_stats.py - the replica-chain statistics of tests/test_gpu_parity_r2.py::test_replica_statistics_headline_geometry
with many more chains on the GPU side (N = 256, main.c geometry, T = A = 1.1: 150 sweeps with 2A, then 100 production sweeps),
to compare with a large reference sample computed on host cores (1024 chains of the oracle's oneParticleMoves restatement,
bit-identical to the compiled reference): <E> = -4.9551 +- 0.0202, acceptance = 0.94930 +- 0.00012, mean height = 16.776 +- 0.055.
The energy and the acceptance agree; the mean height does not, and tests/studies/boxmuller_pairing.py shows why: the
reference's vecBoxMuller couples the two numbers of a pair, the Philox stream of the engine does not."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
smcb = importlib.import_module("montecarlo-surfacer_b200")
from oracle_bindings import GOLDEN_W_M3, Oracle
N, L, LZ, T, A, C = 256, 33.0, 240.0, 1.1, 1.1, 16384
R0, _ = Oracle().initialize_box(L, LZ, N)
with smcb.Engine(C, N, 3) as eng:
    eng.set_params(smcb.default_params(L=L, Lz=LZ, T=T, A=A), GOLDEN_W_M3)
    eng.broadcast_positions(R0)
    eng.set_rng(4711, 0, 0)
    eng.set_step_scale(2.0); eng.sweep(150, smcb.FAST); eng.set_step_scale(1.0)
    eng.reset_counters()
    es = []
    for k in range(10):
        eng.sweep(10, smcb.FAST)
        es.append(eng.chain_state()[0].copy())
    _, na, nt = eng.chain_state()
    gz = eng.get_positions().reshape(C, N, 3)[:, :, 2].mean(axis=1)
E = np.mean(es, axis=0); acc = na / nt
ref = {"<E>": (-4.955065277564973, 0.020182604713564357), "acceptance": (0.9492975616455078, 0.00012183923223226912), "mean height": (16.776123682391074, 0.05490451597908347)}
for name, x in (("<E>", E), ("acceptance", acc), ("mean height", gz)):
    m, s = x.mean(), x.std(ddof=1) / np.sqrt(x.size)
    rm, rs = ref[name]
    print(f"{name:12s} GPU ({C} chains) {m:.5f} +- {s:.5f}   reference (1024 chains) {rm:.5f} +- {rs:.5f}   difference {abs(m - rm) / np.hypot(s, rs):.2f} sigma")
