#!/usr/bin/env python
"""tests/golden/make_golden.py — writes the golden fixtures of this directory.

TEST INFRASTRUCTURE.  Run in the build container, where /root/reference exists and
oracle/build_ref.sh has compiled it into oracle/_ref/libref_N*_M*.so:

    python tests/golden/make_golden.py

Every number below is produced by the UNMODIFIED reference routines (SMC.c:278-351, 557-895,
912-927, 413-465; matematicose.c:183-193) built with `gcc -std=gnu11 -O2 -ffp-contract=off`; the
rand() stream the reference draws is replayed from integers stored in the fixture, so the
fixtures do not depend on glibc's generator.  The reference itself has no tests or golden vectors
(SURVEY.md §4), so these files ARE the pin: the oracle (CPU, `-m "not gpu"`) and the CUDA path
(`-m gpu`) are both compared with them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_bindings import (GOLDEN_W_M3, RAND_MAX, RefLib, RefNoWall, config_droplet, config_gas,  # noqa: E402
                             config_slab, random_walls)

GEOM = {32: (33.0, 200.0), 108: (33.0, 200.0), 256: (33.0, 240.0)}


def static_case(N, M, seed):
    """per-particle and total energies / forces / virials of four configurations"""
    ref = RefLib(N, M)
    L, Lz = GEOM[N]
    rng = np.random.default_rng(seed)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    cfgs = [ref.initializeBox(L, Lz), config_gas(N, L, Lz, rng), config_droplet(N, L, Lz, rng),
            config_slab(N, L, Lz, rng)]
    # one particle pushed outside the slab exercises the dz = +-1e-4 clamp (SMC.c:738-739)
    cfgs[1][2] = Lz / 2 + 0.3
    cfgs[1][5] = -Lz / 2 - 0.2
    out = {"N": N, "M": M, "L": L, "Lz": Lz, "W": W, "R": np.stack(cfgs)}
    keys = ("e_lj", "f_lj", "e_wall", "f_wall", "U_lj", "U_wall", "P_lj", "P_wall", "forces_newton3")
    acc = {k: [] for k in keys}
    for R in cfgs:
        acc["e_lj"].append([ref.energySingle(R, L, i) for i in range(N)])
        acc["f_lj"].append(np.concatenate([ref.forceSingle(R, L, i) for i in range(N)]))
        acc["e_wall"].append([ref.wallsEnergySingle(R[3 * i:3 * i + 3], W, L, Lz) for i in range(N)])
        acc["f_wall"].append(np.concatenate([ref.wallsForce(R[3 * i:3 * i + 3], W, L, Lz) for i in range(N)]))
        acc["U_lj"].append(ref.energy(R, L))
        acc["U_wall"].append(ref.wallsEnergy(R, W, L, Lz))
        acc["P_lj"].append(ref.pressure(R, L, Lz))
        acc["P_wall"].append(ref.wallsPressure(R, W, L, Lz))
        acc["forces_newton3"].append(ref.forces(R, L, np.zeros(3 * N)))
    out.update({k: np.array(v) for k, v in acc.items()})
    return out


def sweep_case(N, M, T, A, start, nsweeps, seed):
    """oneParticleMoves driven by a replayed rand() stream: state after every sweep"""
    ref = RefLib(N, M)
    L, Lz = GEOM[N]
    rng = np.random.default_rng(seed)
    W = GOLDEN_W_M3.copy()
    R = {"lattice": lambda: ref.initializeBox(L, Lz), "droplet": lambda: config_droplet(N, L, Lz, rng, jitter=0.03),
         "slab": lambda: config_slab(N, L, Lz, rng)}[start]()
    R0 = R.copy()
    per = 4 * N + 1
    stream = rng.integers(0, RAND_MAX, size=per * nsweeps, endpoint=True, dtype=np.int64).astype(np.int32)
    ref.set_replay(stream)
    E = ref.energy(R, L) + ref.wallsEnergy(R, W, L, Lz)
    E0 = E
    Es, js, Rs = [], [], []
    for _ in range(nsweeps):
        j, E = ref.oneParticleMoves(R, W, L, Lz, A, T, E)
        Es.append(E)
        js.append(j)
        Rs.append(R.copy())
    assert ref.replay_pos() == per * nsweeps and ref.replay_underflow() == 0
    ref.set_replay(None)
    keep = sorted(set([0, 1, nsweeps // 2, nsweeps - 1]))
    return {"N": N, "M": M, "L": L, "Lz": Lz, "T": T, "A": A, "W": W, "R0": R0, "E0": E0, "stream": stream.reshape(nsweeps, per),
            "E": np.array(Es), "naccept": np.array(js), "R_at": np.array(keep), "R": np.stack([Rs[k] for k in keep])}


def misc_case():
    """vecBoxMuller from a replayed stream, initializeBox lattices, localDensityAndMobility"""
    ref = RefLib(108, 3)
    rng = np.random.default_rng(99)
    ints = rng.integers(0, RAND_MAX, size=324, endpoint=True, dtype=np.int64).astype(np.int32)
    ref.set_replay(ints)
    bm = ref.vecBoxMuller(np.sqrt(2 * 1.1), 324)
    ref.set_replay(None)
    L, Lz = GEOM[108]
    D = np.zeros(33 ** 3, dtype=np.uint64)
    Mu = np.zeros(33 ** 3, dtype=np.uint64)
    Rbin = np.zeros(108, dtype=np.int32)
    Ra = config_gas(108, L, Lz, rng)
    Rb = Ra + 0.8 * rng.standard_normal(324)
    Rb[0::3] -= L * np.rint(Rb[0::3] / L)
    Rb[1::3] -= L * np.rint(Rb[1::3] / L)
    ref.localDensityAndMobility(Ra, L, Lz, D, Rbin, Mu)
    ref.localDensityAndMobility(Rb, L, Lz, D, Rbin, Mu)
    nz = np.flatnonzero(D)
    nzm = np.flatnonzero(Mu)
    out = {"bm_ints": ints, "bm_sigma": np.sqrt(2 * 1.1), "bm_out": bm,
           "ld_Ra": Ra, "ld_Rb": Rb, "ld_D_idx": nz, "ld_D_val": D[nz], "ld_Mu_idx": nzm, "ld_Mu_val": Mu[nzm], "ld_Rbin": Rbin,
           "box108": RefLib(108, 3).initializeBox(33.0, 200.0), "box256": RefLib(256, 3).initializeBox(33.0, 240.0),
           "box32": RefLib(32, 3).initializeBox(33.0, 200.0)}
    # bulk prototype (SMC_noMPI_noWall.c:464-493, 573-591, 664-684): 3-D minimum image, cutoff L/2
    nw = RefNoWall(108)
    Lb = (108 / 0.5) ** (1.0 / 3.0)
    Rw = nw.initializeBox(Lb) + 0.05 * (rng.random(324) * 2 - 1)
    out.update(bulk_L=Lb, bulk_R=Rw, bulk_energy=nw.energy(Rw, Lb), bulk_forces=nw.forces(Rw, Lb), bulk_pressure=nw.pressure(Rw, Lb))
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "static_N32_M3.npz"), **static_case(32, 3, 1))
    np.savez_compressed(os.path.join(HERE, "static_N108_M3.npz"), **static_case(108, 3, 2))
    np.savez_compressed(os.path.join(HERE, "static_N256_M3.npz"), **static_case(256, 3, 3))
    np.savez_compressed(os.path.join(HERE, "static_N108_M4.npz"), **static_case(108, 4, 4))
    np.savez_compressed(os.path.join(HERE, "sweep_N108_lattice.npz"), **sweep_case(108, 3, 1.1, 1.1, "lattice", 24, 11))
    np.savez_compressed(os.path.join(HERE, "sweep_N108_droplet.npz"), **sweep_case(108, 3, 0.8, 0.004, "droplet", 16, 12))
    np.savez_compressed(os.path.join(HERE, "sweep_N32_slab.npz"), **sweep_case(32, 3, 1.1, 1.1, "slab", 40, 13))
    np.savez_compressed(os.path.join(HERE, "sweep_N256_droplet.npz"), **sweep_case(256, 3, 1.1, 0.01, "droplet", 6, 14))
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **misc_case())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
