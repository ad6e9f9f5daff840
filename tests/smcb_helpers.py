"""shared helpers of the GPU parity tests"""
import importlib

import numpy as np

from oracle_bindings import (GOLDEN_W_M3, RAND_MAX, Oracle, config_droplet, config_gas, config_slab, make_sys,
                             random_walls)

smcb = importlib.import_module("montecarlo-surfacer_b200")

GEOM = {32: (33.0, 200.0), 64: (33.0, 200.0), 108: (33.0, 200.0), 256: (33.0, 240.0), 500: (33.0, 240.0),
        4096: (33.0, 240.0)}


def geom(N):
    return GEOM.get(N, (33.0, 240.0 if N >= 150 else 200.0))


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor-scale): relative error with an absolute floor for values near 0"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), floor)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def mixed_configs(N, L, Lz, nchains, seed, orc):
    """nchains different configurations: lattice, gas, droplet, slab, repeated with new seeds"""
    rng = np.random.default_rng(seed)
    out = []
    kinds = ["lattice", "gas", "droplet", "slab"] if N <= 256 else ["gas", "droplet"]
    for c in range(nchains):
        kind = kinds[c % len(kinds)]
        if kind == "lattice":
            X, sites = orc.initialize_box(L, Lz, N)
            if sites != N:
                X = config_gas(N, L, Lz, rng)
            else:
                X = X + (rng.random(3 * N) - 0.5) * 0.01 * (c > 0)
            out.append(X)
        elif kind == "gas":
            out.append(config_gas(N, L, Lz, rng))
        elif kind == "droplet":
            out.append(config_droplet(N, L, Lz, rng, jitter=0.03 + 0.04 * rng.random(), nz=4 if N <= 500 else 8))
        else:
            out.append(config_slab(N, L, Lz, rng))
    return np.stack(out)


def make_stream(N, nsweeps, rng):
    """the 4N+1 rand() integers per sweep the reference would draw"""
    per = 4 * N + 1
    return rng.integers(0, RAND_MAX, size=per * nsweeps, endpoint=True, dtype=np.int64).astype(np.int32).reshape(nsweeps, per)


def expand_streams(orc, N, A, streams):
    """streams [S, C, 4N+1] ints -> displ [S,C,3N], offset [S,C], u [S,C,N] as the reference derives them"""
    S, Cn, _ = streams.shape
    displ = np.empty((S, Cn, 3 * N))
    off = np.empty((S, Cn), dtype=np.int64)
    u = np.empty((S, Cn, N))
    for s in range(S):
        for c in range(Cn):
            displ[s, c], off[s, c], u[s, c] = orc.expand_stream(N, A[c] if np.ndim(A) else A, streams[s, c])
    return displ, off, u
