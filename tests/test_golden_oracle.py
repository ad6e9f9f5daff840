"""The oracle (oracle/smc_oracle.c) against the committed golden fixtures (tests/golden/*.npz),
which tests/golden/make_golden.py generated from the UNMODIFIED reference compiled from
/root/reference.  Unlike test_oracle_vs_ref.py these need no oracle/_ref, so they pin the oracle on
any host.  Bit-for-bit (both sides built -O2 -ffp-contract=off)."""
import glob
import os

import numpy as np
import pytest

from oracle_bindings import Oracle, make_sys

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_fixture_set_is_complete():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    assert {"static_N32_M3.npz", "static_N108_M3.npz", "static_N256_M3.npz", "static_N108_M4.npz",
            "sweep_N108_lattice.npz", "sweep_N108_droplet.npz", "sweep_N32_slab.npz", "sweep_N256_droplet.npz",
            "misc.npz"} <= names


@pytest.mark.parametrize("name", ["static_N32_M3.npz", "static_N108_M3.npz", "static_N256_M3.npz", "static_N108_M4.npz"])
def test_static_golden(orc, name):
    g = _load(name)
    N, M, L, Lz, W = int(g["N"]), int(g["M"]), float(g["L"]), float(g["Lz"]), g["W"]
    s = make_sys(N, M, L, Lz)
    for c, R in enumerate(g["R"]):
        R = np.ascontiguousarray(R)
        assert orc.energy(s, R) == g["U_lj"][c]
        assert orc.walls_energy(s, R, W) == g["U_wall"][c]
        assert orc.pressure(s, R) == g["P_lj"][c]
        assert orc.walls_pressure(s, R, W) == g["P_wall"][c]
        np.testing.assert_array_equal(orc.forces(s, R), g["forces_newton3"][c])
        np.testing.assert_array_equal([orc.energy_single(s, R, i) for i in range(N)], g["e_lj"][c])
        np.testing.assert_array_equal(np.concatenate([orc.force_single(s, R, i) for i in range(N)]), g["f_lj"][c])
        np.testing.assert_array_equal([orc.walls_energy_single(s, R[3 * i:3 * i + 3], W) for i in range(N)], g["e_wall"][c])
        np.testing.assert_array_equal(np.concatenate([orc.walls_force(s, R[3 * i:3 * i + 3], W) for i in range(N)]), g["f_wall"][c])


@pytest.mark.parametrize("name", ["sweep_N108_lattice.npz", "sweep_N108_droplet.npz", "sweep_N32_slab.npz", "sweep_N256_droplet.npz"])
def test_sweep_golden(orc, name):
    g = _load(name)
    N, M, L, Lz = int(g["N"]), int(g["M"]), float(g["L"]), float(g["Lz"])
    s = make_sys(N, M, L, Lz)
    R = g["R0"].copy()
    E = orc.energy(s, R) + orc.walls_energy(s, R, g["W"])
    assert E == float(g["E0"])
    keep = {int(k): i for i, k in enumerate(g["R_at"])}
    for k, ints in enumerate(g["stream"]):
        j, E = orc.sweep_from_ints(s, R, g["W"], float(g["A"]), float(g["T"]), ints, E)
        assert j == g["naccept"][k] and E == g["E"][k], k
        if k in keep:
            np.testing.assert_array_equal(R, g["R"][keep[k]], err_msg=f"sweep {k}")


def test_misc_golden(orc):
    g = _load("misc.npz")
    np.testing.assert_array_equal(orc.box_muller(float(g["bm_sigma"]), 324, g["bm_ints"]), g["bm_out"])
    for n, key, Lz in ((108, "box108", 200.0), (256, "box256", 240.0), (32, "box32", 200.0)):
        X, sites = orc.initialize_box(33.0, Lz, n)
        np.testing.assert_array_equal(X, g[key])
    s = make_sys(108, 3, 33.0, 200.0)
    D = np.zeros(33 ** 3, dtype=np.uint64)
    Mu = np.zeros(33 ** 3, dtype=np.uint64)
    Rbin = np.zeros(108, dtype=np.int32)
    orc.local_density(s, np.ascontiguousarray(g["ld_Ra"]), D, Rbin, Mu)
    orc.local_density(s, np.ascontiguousarray(g["ld_Rb"]), D, Rbin, Mu)
    np.testing.assert_array_equal(np.flatnonzero(D), g["ld_D_idx"])
    np.testing.assert_array_equal(D[g["ld_D_idx"]], g["ld_D_val"])
    np.testing.assert_array_equal(np.flatnonzero(Mu), g["ld_Mu_idx"])
    np.testing.assert_array_equal(Mu[g["ld_Mu_idx"]], g["ld_Mu_val"])
    np.testing.assert_array_equal(Rbin, g["ld_Rbin"])
    Lb = float(g["bulk_L"])
    sb = make_sys(108, 3, Lb, Lb, rc2=Lb * Lb / 4, periodic_z=1, wall=0)
    Rw = np.ascontiguousarray(g["bulk_R"])
    assert orc.energy(sb, Rw) == float(g["bulk_energy"])
    # the prototype's pressure() walks the pairs l-major (SMC_noMPI_noWall.c:668-669), SMC.c's i-major
    # (SMC.c:702-703, the order the oracle keeps): same terms, different summation order -> last-ulp only
    assert abs(orc.pressure(sb, Rw) - float(g["bulk_pressure"])) <= 4e-16 * abs(float(g["bulk_pressure"]))
    np.testing.assert_array_equal(orc.forces(sb, Rw), g["bulk_forces"])
