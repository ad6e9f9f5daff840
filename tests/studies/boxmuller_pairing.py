#!/usr/bin/env python
"""tests/studies/boxmuller_pairing.py - why long replica runs of the reference drift upwards and the engine's do not.

vecBoxMuller (matematicose.c:183-193) fills A[2i] = s*sqrt(-2 ln(1-x1))*cos(2 pi x2) and A[2i+1] = s*sqrt(-2 ln(1-x2))*sin(2 pi x1):
the radius of one output is drawn from the angle variable of the other.  Each output is an exact Gaussian, but the two are
NOT independent (E[A0 A1]/s^2 = -0.0174, and the higher mixed moments do not vanish either), and in displ[3N] the pairs
straddle (x,y), (z, x of the next molecule) or (y,z) of a molecule depending on the parity of 3n.  The Metropolis
test of oneParticleMoves (SMC.c:331-335) assumes an isotropic proposal, so with this stream detailed balance holds only
approximately.  The visible effect in main.c's geometry (N = 256, T = A = 1.1, 150 sweeps at 2A + 100 sweeps): the mean
height of the gas drifts from 16.50 to 16.78 +- 0.05 with the reference's stream and stays at 16.50 with independent
Gaussians - same code, same acceptance rule, only the joint law of the displacements differs.

  CPU part (no GPU needed, ~1 min on 8 cores): the oracle's sweep driven by (a) numpy's independent normals,
      (b) the reference's mapping of 4N+1 integers (expand_stream), 384 chains each.
  GPU part (--gpu): the engine's FED sweeps (the reference's acceptance rule, FAST arithmetic) on 4096 chains with the
      same two kinds of stream built in numpy, and its own Philox stream.
This is a study, not a test: it is not collected by pytest."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_bindings import GOLDEN_W_M3, Oracle, make_sys, RAND_MAX

N, M, T, A, L, LZ = 256, 3, 1.1, 1.1, 33.0, 240.0
NPRE, NPROD = 150, 100


def pair_moments(n=4_000_000, seed=1):
    rng = np.random.default_rng(seed)
    x1, x2 = rng.random(n), rng.random(n)
    a0 = np.sqrt(-2 * np.log(1 - x1)) * np.cos(2 * np.pi * x2)
    a1 = np.sqrt(-2 * np.log(1 - x2)) * np.sin(2 * np.pi * x1)
    return {"E[a0 a1]": np.mean(a0 * a1), "E[a0^2 a1]": np.mean(a0 * a0 * a1), "E[a0 a1^2]": np.mean(a0 * a1 * a1),
            "E[|a0| a1]": np.mean(np.abs(a0) * a1), "E[a0 |a1|]": np.mean(a0 * np.abs(a1)), "err": 1 / np.sqrt(n)}


def cpu_chain(args):
    c, mode = args
    orc = Oracle(); s = make_sys(N, M, L, LZ); W = GOLDEN_W_M3.copy()
    R, _ = orc.initialize_box(L, LZ, N); E = 0.0
    rng = np.random.default_rng(10_000 * mode + c)
    for k in range(NPRE + NPROD):
        Ak = 2 * A if k < NPRE else A
        if mode == 0:
            displ = rng.standard_normal(3 * N) * np.sqrt(2 * Ak); off = int(rng.integers(0, 2**31)); u = rng.random(N)
        else:
            ints = rng.integers(0, RAND_MAX, size=4 * N + 1, endpoint=True).astype(np.int32)
            displ, off, u = orc.expand_stream(N, Ak, ints)
        _, E = orc.sweep(s, R, W, Ak, T, displ, off, u, E)
    return R[2::3].mean()


def gpu_part(C=4096, batch=5):
    smcb = importlib.import_module("montecarlo-surfacer_b200")
    R0, _ = Oracle().initialize_box(L, LZ, N)
    out = {}
    for kind in ("independent", "paired", "philox"):
        rng = np.random.default_rng(77)
        with smcb.Engine(C, N, M) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=LZ, T=T, A=A), GOLDEN_W_M3)
            eng.broadcast_positions(R0)
            eng.set_rng(4711, 0, 0)
            for k0 in range(0, NPRE + NPROD, batch):
                scale = 2.0 if k0 < NPRE else 1.0
                eng.set_step_scale(scale)
                if kind == "philox":
                    eng.sweep(batch, smcb.FAST)
                    continue
                sg = np.sqrt(2 * A * scale)
                if kind == "independent":
                    displ = rng.standard_normal((batch, C, 3 * N)) * sg
                else:
                    x1, x2 = rng.random((batch, C, 3 * N // 2)), rng.random((batch, C, 3 * N // 2))
                    displ = np.empty((batch, C, 3 * N))
                    displ[..., 0::2] = sg * np.sqrt(-2 * np.log(1 - x1)) * np.cos(2 * np.pi * x2)
                    displ[..., 1::2] = sg * np.sqrt(-2 * np.log(1 - x2)) * np.sin(2 * np.pi * x1)
                eng.sweep_fed(displ, rng.integers(0, 2**31, size=(batch, C)), rng.random((batch, C, N)), smcb.FAST)
            E, na, nt = eng.chain_state()
            z = eng.get_positions().reshape(C, N, 3)[:, :, 2].mean(axis=1)
        out[kind] = (z.mean(), z.std(ddof=1) / np.sqrt(C), E.mean(), E.std(ddof=1) / np.sqrt(C), (na / nt).mean())
        print(f"GPU fed sweeps, {kind:12s} stream ({C} chains): mean height {out[kind][0]:.4f} +- {out[kind][1]:.4f}   "
              f"E {out[kind][2]:.4f} +- {out[kind][3]:.4f}   acceptance {out[kind][4]:.5f}", flush=True)
    return out


if __name__ == "__main__":
    print("mixed moments of one vecBoxMuller pair (sigma = 1):", {k: round(float(v), 5) for k, v in pair_moments().items()})
    if "--gpu" in sys.argv:
        gpu_part()
    else:
        from multiprocessing import Pool
        with Pool(min(8, os.cpu_count() or 1)) as p:
            for mode, nm in ((0, "independent Gaussians"), (1, "reference pairing")):
                z = np.array(p.map(cpu_chain, [(c, mode) for c in range(384)]))
                print(f"oracle sweep, {nm:22s}: mean height {z.mean():.4f} +- {z.std(ddof=1) / np.sqrt(z.size):.4f}")
