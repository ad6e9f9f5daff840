#!/usr/bin/env python
"""profiles/spec_stats.py - how the trials of k_sweep_spec split over its paths (library built with
`make -C montecarlo-surfacer_b200/csrc EXTRA=-DSMCB_SPEC_STATS`), lattice start and thermalised gas."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
smcb = importlib.import_module("montecarlo-surfacer_b200")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
N, C, S, L, Lz, T = 256, 2048, 40, 33.0, 240.0, 1.1
a = L / 4
cells = np.array([(i, j, k) for i in range(4) for j in range(4) for k in range(4)], dtype=float)
basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
R0 = ((cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + a / 4).reshape(-1)
names = ["trials", "lonely parallel commits", "serial commits", "general path", "  general: dense segment", "  general: several partners",
         "  general: near the surface", "  general: dirty caches", "valid rejections (hard, nothing to do)", "serial commits with old partners",
         "serial commits with a new partner", "trials whose molecule has partners"]
with smcb.Engine(C, N, 3) as eng:
    eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=T), smcb.REFERENCE_WALL_M3)
    eng.broadcast_positions(R0); eng.set_rng(12345, 0, 0)
    for label, pre in (("lattice start (sweeps 120-160)", 120), ("after 2000 sweeps with 2A", 2000)):
        if pre == 2000:
            eng.set_step_scale(2.0); eng.sweep(2000, smcb.FAST); eng.set_step_scale(1.0)
        else:
            eng.sweep(pre, smcb.FAST)
        eng.sweep(S, smcb.FAST)
        st = eng.debug_sweep_stats().astype(float)
        print(label)
        for nm, v in zip(names, st):
            print(f"  {nm:42s} {v / st[0]:8.4f} per trial")
        print(f"  {'general: conflict with an accepted trial':42s} {(st[3] - st[4] - st[5] - st[6] - st[7]) / st[0]:8.4f} per trial")
