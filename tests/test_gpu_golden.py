"""GPU parity against the committed golden fixtures (tests/golden/*.npz, generated from the
UNMODIFIED reference by tests/golden/make_golden.py), through the C ABI.
  STRICT: per-particle routines and whole sweep trajectories BIT-IDENTICAL to the reference
          (SMC.c:278-351, 557-618, 729-813); chain totals within 1e-12 (different summation tree).
  FAST  : static values within 1e-12 relative; sweeps teacher-forced per sweep within 1e-12."""
import os

import numpy as np
import pytest

from smcb_helpers import Oracle, rel_err, smcb

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


@pytest.mark.parametrize("name", ["static_N32_M3.npz", "static_N108_M3.npz", "static_N256_M3.npz", "static_N108_M4.npz"])
def test_static_golden_gpu(name):
    g = np.load(os.path.join(GOLD, name))
    N, M, L, Lz, W = int(g["N"]), int(g["M"]), float(g["L"]), float(g["Lz"]), g["W"]
    R = g["R"]
    C = R.shape[0]
    vol3 = 3 * L * L * Lz
    with smcb.Engine(C, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz), W)
        eng.set_positions(R)
        strict = eng.evaluate(smcb.STRICT)
        fast = eng.evaluate(smcb.FAST)
    for k in ("e_lj", "f_lj", "e_wall", "f_wall"):
        np.testing.assert_array_equal(strict[k], g[k], err_msg=f"STRICT {k}")
        for c in range(C):
            scale = max(1.0, float(np.max(np.abs(g[k][c]))) * 1e-3)
            assert rel_err(fast[k][c], g[k][c], floor=scale) < TOL, (k, c)
    # forces() (Newton-3 accumulation, SMC.c:656-686) equals forceSingle per particle to rounding
    for c in range(C):
        scale = max(1.0, float(np.max(np.abs(g["forces_newton3"][c]))) * 1e-3)
        assert rel_err(strict["f_lj"][c], g["forces_newton3"][c], floor=scale) < 1e-11
    for res in (strict, fast):
        for c in range(C):
            for k, ref in (("U_lj", g["U_lj"][c]), ("U_wall", g["U_wall"][c]),
                           ("vir_lj", -g["P_lj"][c] * vol3), ("vir_wall_ref", -g["P_wall"][c] * vol3)):
                assert abs(res[k][c] - ref) <= TOL * max(1.0, abs(ref)), (k, c, res[k][c], ref)


@pytest.mark.parametrize("name", ["sweep_N108_lattice.npz", "sweep_N108_droplet.npz", "sweep_N32_slab.npz", "sweep_N256_droplet.npz"])
def test_sweep_golden_gpu_strict(name):
    """the reference's rand() integers are expanded exactly as it does (matematicose.c:183-193,
    SMC.c:290,335) and fed to the STRICT sweep kernel: every stored state must match bit for bit"""
    g = np.load(os.path.join(GOLD, name))
    N, M, L, Lz, T, A = int(g["N"]), int(g["M"]), float(g["L"]), float(g["Lz"]), float(g["T"]), float(g["A"])
    orc = Oracle()
    S = g["stream"].shape[0]
    displ = np.empty((S, 1, 3 * N)); off = np.empty((S, 1), dtype=np.int64); u = np.empty((S, 1, N))
    for k in range(S):
        displ[k, 0], off[k, 0], u[k, 0] = orc.expand_stream(N, A, g["stream"][k])
    keep = {int(k): i for i, k in enumerate(g["R_at"])}
    with smcb.Engine(1, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), g["W"])
        eng.set_positions(g["R0"])
        eng.refresh_energy(smcb.STRICT)
        E0 = float(g["E0"])
        assert abs(eng.chain_state()[0][0] - E0) <= TOL * max(1.0, abs(E0))   # chain total: tree sum, 1e-12
        eng.set_chain_energy([E0])                                           # sMC seeds E[0] itself (SMC.c:48)
        nacc_prev = 0
        for k in range(S):
            eng.sweep_fed(displ[k:k + 1], off[k:k + 1], u[k:k + 1], mode=smcb.STRICT)
            E, na, nt = eng.chain_state()
            assert E[0] == g["E"][k], (k, E[0], g["E"][k])
            assert na[0] - nacc_prev == g["naccept"][k], k
            nacc_prev = na[0]
            if k in keep:
                np.testing.assert_array_equal(eng.get_positions()[0], g["R"][keep[k]], err_msg=f"sweep {k}")


@pytest.mark.parametrize("name", ["sweep_N108_lattice.npz", "sweep_N32_slab.npz"])
def test_sweep_golden_gpu_fast_first_sweep(name):
    """FAST kernel, first golden sweep from the golden start: same accept count, energy within 1e-12
    (later sweeps diverge chaotically from 1e-16 differences - SURVEY §0-5 - and are covered by the
    teacher-forced tests in test_gpu_sweep.py)"""
    g = np.load(os.path.join(GOLD, name))
    N, M, L, Lz, T, A = int(g["N"]), int(g["M"]), float(g["L"]), float(g["Lz"]), float(g["T"]), float(g["A"])
    orc = Oracle()
    displ, off, u = orc.expand_stream(N, A, g["stream"][0])
    with smcb.Engine(1, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), g["W"])
        eng.set_positions(g["R0"])
        eng.sweep_fed(displ[None, None], np.array([[off]], dtype=np.int64), u[None, None], mode=smcb.FAST)
        E, na, _ = eng.chain_state()
        assert na[0] == g["naccept"][0]
        assert abs(E[0] - g["E"][0]) <= TOL * max(1.0, abs(g["E"][0]))
        if 0 in list(g["R_at"]):
            assert rel_err(eng.get_positions()[0], g["R"][0], floor=1.0) < TOL


def test_local_density_golden_gpu():
    g = np.load(os.path.join(GOLD, "misc.npz"))
    with smcb.Engine(1, 108, 3) as eng:
        eng.set_params(smcb.default_params(L=33.0, Lz=200.0), np.zeros(18))
        for key in ("ld_Ra", "ld_Rb"):
            eng.set_positions(g[key])
            eng.gather()
        o = eng.obs_get()[0]
        D, Mu = o["D"].reshape(-1), o["Mu"].reshape(-1)
        np.testing.assert_array_equal(np.flatnonzero(D), g["ld_D_idx"])
        np.testing.assert_array_equal(D[g["ld_D_idx"]], g["ld_D_val"])
        np.testing.assert_array_equal(np.flatnonzero(Mu), g["ld_Mu_idx"])
        np.testing.assert_array_equal(Mu[g["ld_Mu_idx"]], g["ld_Mu_val"])
        np.testing.assert_array_equal(eng.rbin()[0], g["ld_Rbin"])


def test_bulk_golden_gpu():
    """config 1 geometry: the bulk prototype's energy/forces/pressure (SMC_noMPI_noWall.c:464-493,
    573-591, 664-684) with PERIODIC_Z"""
    g = np.load(os.path.join(GOLD, "misc.npz"))
    Lb = float(g["bulk_L"])
    with smcb.Engine(1, 108, 3) as eng:
        eng.set_params(smcb.default_params(L=Lb, Lz=Lb, rc2=Lb * Lb / 4, flags=smcb.PERIODIC_Z))
        eng.set_positions(g["bulk_R"])
        for mode in (smcb.STRICT, smcb.FAST):
            ev = eng.evaluate(mode)
            assert abs(ev["U_lj"][0] - float(g["bulk_energy"])) <= TOL * abs(float(g["bulk_energy"]))
            assert abs(-ev["vir_lj"][0] / (3 * Lb ** 3) - float(g["bulk_pressure"])) <= TOL * abs(float(g["bulk_pressure"]))
            scale = float(np.max(np.abs(g["bulk_forces"]))) * 1e-3
            assert rel_err(ev["f_lj"][0], g["bulk_forces"], floor=scale) < 1e-11
