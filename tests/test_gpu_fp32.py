"""The optional single-precision mode (SMCB_FP32, csrc/fp32_mode.cuh): north_star's "1e-5 relative in the optional
fp32 mode", against the same oracle as the FP64 kernels (the reference's energySingle / forceSingle / wallsEnergySingle
/ wallsForce / energy / pressure, SMC.c:557-895).  "Relative" is taken per element against the magnitude of what the
element sums (|f| or the sum of the |pair terms| of that component), as in test_gpu_parity_r2.py."""
import numpy as np
import pytest

from smcb_helpers import GOLDEN_W_M3, Oracle, geom, make_sys, mixed_configs, smcb

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _pair_scales(R, Lbox):
    """per molecule: sum_j |4 e_ij| and per component sum_j |g_ij d_ij,c|"""
    X = R.reshape(-1, 3)
    d = X[:, None, :] - X[None, :, :]
    d[:, :, 0] -= Lbox * np.rint(d[:, :, 0] / Lbox)
    d[:, :, 1] -= Lbox * np.rint(d[:, :, 1] / Lbox)
    r2 = np.einsum("ijk,ijk->ij", d, d)
    np.fill_diagonal(r2, np.inf)
    inside = r2 < 9.0
    r2 = np.where(inside, r2, 1.0)
    e = np.where(inside, 4.0 * (np.abs(1.0 / r2 ** 6) + np.abs(1.0 / r2 ** 3)), 0.0)
    g = np.where(inside, np.abs(48.0 / r2 ** 7) + np.abs(24.0 / r2 ** 4), 0.0)
    return e.sum(axis=1), np.sum(g[:, :, None] * np.abs(d), axis=1)


def _wall_scales(R, W, Lbox, Lz, M=3, a0=smcb.engine.A0_DEFAULT, b0=smcb.engine.B0_DEFAULT):
    """per molecule: sum of the magnitudes of the surface terms of its energy (x4) and of each force component
    (flat wall a0/dz^12 - b0/dz^6 and the M x M sites, SMC.c:729-813)"""
    X = R.reshape(-1, 3)
    dz = X[:, 2] + Lz / 2
    dz = dz - Lz * np.rint(dz / Lz)
    dz = np.where(X[:, 2] <= -Lz / 2, 1e-4, np.where(X[:, 2] >= Lz / 2, -1e-4, dz))
    We = 4.0 * (a0 / dz ** 12 + b0 / dz ** 6)
    Wf = np.zeros_like(X)
    Wf[:, 2] = (48.0 * a0 / dz ** 14 + 24.0 * b0 / dz ** 8) * np.abs(dz)
    dw = Lbox / M
    for i in range(M):
        for j in range(M):
            m = j + i * M
            dx = X[:, 0] - i * dw
            dx = dx - Lbox * np.rint(dx / Lbox)
            dy = X[:, 1] - j * dw
            dy = dy - Lbox * np.rint(dy / Lbox)
            r2 = dx * dx + dy * dy + dz * dz
            inside = r2 < 9.0
            r2 = np.where(inside, r2, 1.0)
            We += np.where(inside, 4.0 * (W[2 * m] / r2 ** 6 + W[2 * m + 1] / r2 ** 3), 0.0)
            g = np.where(inside, 48.0 * W[2 * m] / r2 ** 7 + 24.0 * W[2 * m + 1] / r2 ** 4, 0.0)
            Wf += g[:, None] * np.abs(np.stack([dx, dy, dz], axis=1))
    return We, Wf


@pytest.mark.parametrize("n", [256, 108, 33, 500])
def test_fp32_evaluation_within_1e5(orc, n):
    L, Lz = geom(n)
    s = make_sys(n, 3, L, Lz)
    W = GOLDEN_W_M3.copy()
    R = mixed_configs(n, L, Lz, 6, seed=7 * n, orc=orc)
    with smcb.Engine(R.shape[0], n, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz), W)
        eng.set_positions(R)
        ev = eng.evaluate(smcb.FP32)
    for c in range(R.shape[0]):
        Se, Sf = _pair_scales(R[c], L)
        e_lj = np.array([orc.energy_single(s, R[c], i) for i in range(n)])
        f_lj = np.concatenate([orc.force_single(s, R[c], i) for i in range(n)]).reshape(n, 3)
        e_w = np.array([orc.walls_energy_single(s, R[c][3 * i:3 * i + 3], W) for i in range(n)])
        f_w = np.concatenate([orc.walls_force(s, R[c][3 * i:3 * i + 3], W) for i in range(n)]).reshape(n, 3)
        assert np.all(np.abs(ev["e_lj"][c] - e_lj) <= TOL * np.maximum(np.abs(e_lj), Se) + 1e-30), c
        assert np.all(np.abs(ev["f_lj"][c].reshape(n, 3) - f_lj) <= TOL * np.maximum(np.abs(f_lj), Sf) + 1e-30), c
        # surface terms: the flat wall and the sites in range, each a difference of a repulsive and an attractive part
        We, Wf = _wall_scales(R[c], W, L, Lz)
        assert np.all(np.abs(ev["e_wall"][c] - e_w) <= TOL * np.maximum(np.abs(e_w), We) + 1e-30), c
        assert np.all(np.abs(ev["f_wall"][c].reshape(n, 3) - f_w) <= TOL * np.maximum(np.abs(f_w), Wf) + 1e-30), c
        U, Uw, P = orc.energy(s, R[c]), orc.walls_energy(s, R[c], W), orc.pressure(s, R[c])
        assert abs(ev["U_lj"][c] - U) <= TOL * max(abs(U), 0.5 * Se.sum())
        assert abs(ev["U_wall"][c] - Uw) <= TOL * max(abs(Uw), We.sum())
        vol3 = 3 * L * L * Lz
        assert abs(-ev["vir_lj"][c] / vol3 - P) <= TOL * max(abs(P), 6 * Se.sum() / vol3)


@pytest.mark.parametrize("n,Astep", [(256, 1e-4), (108, 2e-4)])
def test_fp32_allparticle_step_teacher_forced(orc, n, Astep):
    """one FP32 all-particle step from the oracle's state at a time: ln ap within 1e-5 of the magnitude of what it
    sums, the same accept decision unless log u falls inside that margin, new positions to 1e-6"""
    L, Lz = geom(n)
    T = 1.1
    s = make_sys(n, 3, L, Lz)
    W = GOLDEN_W_M3.copy()
    C, nsteps = 4, 10
    rng = np.random.default_rng(n)
    R = mixed_configs(n, L, Lz, C, seed=3 * n, orc=orc)
    nacc = 0
    with smcb.Engine(C, n, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=Astep), W)
        for k in range(nsteps):
            xi = rng.standard_normal((1, C, 3 * n)) * np.sqrt(2 * Astep)
            u = rng.random((1, C))
            eng.set_positions(R)
            lnap, acc = eng.step_allparticle_fed(xi, u, mode=smcb.FP32)
            Rg = eng.get_positions()
            for c in range(C):
                F, Ulj, Uw, _ = orc.total(s, R[c], W)
                U = Ulj + Uw
                d = F * (Astep / T) + xi[0, c]
                Rp = R[c] + d
                Rp[0::3] -= L * np.rint(Rp[0::3] / L)
                Rp[1::3] -= L * np.rint(Rp[1::3] / L)
                Fp, Uljp, Uwp, _ = orc.total(s, Rp, W)
                Se, _ = _pair_scales(R[c], L)
                S = (Se.sum() + abs(Uw) + abs(Uwp) + 0.5 * np.sum(np.abs(d * (Fp + F))) + Astep / (4 * T) * np.sum(Fp * Fp + F * F)) / T
                Rold = R[c].copy()
                ok, Unew, ln = orc.allparticle_step(s, R[c], F, U, W, Astep, T, xi[0, c], u[0, c])
                assert abs(lnap[0, c] - ln) <= TOL * max(1.0, S), (k, c, lnap[0, c], ln, S)
                if abs(np.log(u[0, c]) - ln) > 2 * TOL * max(1.0, S):
                    assert bool(acc[0, c]) == bool(ok)
                    assert np.max(np.abs(Rg[c] - R[c])) < 1e-6
                else:                                   # a decision inside the FP32 margin: follow the kernel
                    R[c] = Rg[c] if acc[0, c] else Rold
                nacc += int(acc[0, c])
    assert nacc > 0


def test_fp32_allparticle_free_run_and_mode_limits(orc):
    """free-running FP32 steps from the Philox stream: the carried energy equals an FP32 re-evaluation to 1e-5 of its
    scale, acceptance is sane; the sweep refuses the mode"""
    n = 256
    L, Lz = geom(n)
    R0, _ = orc.initialize_box(L, Lz, n)
    with smcb.Engine(64, n, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=2e-4), GOLDEN_W_M3)
        eng.broadcast_positions(R0)
        eng.set_rng(3, 0, 0)
        eng.step_allparticle(50, smcb.FP32)
        E, na, nt = eng.chain_state()
        ev32 = eng.evaluate(smcb.FP32, per_particle=False)
        ev64 = eng.evaluate(smcb.FAST, per_particle=False)
        assert np.all(nt == 50) and na.sum() > 0.5 * nt.sum()
        for ev in (ev32, ev64):
            Erec = ev["U_lj"] + ev["U_wall"]
            assert np.all(np.abs(E - Erec) <= 1e-5 * np.maximum(1.0, np.abs(Erec)))
        eng.step_allparticle(5, smcb.FAST)            # an FP64 step afterwards refreshes the forces the FP32 step left as floats
        E2 = eng.chain_state()[0]
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        assert np.all(np.abs(E2 - (ev["U_lj"] + ev["U_wall"])) <= 1e-9 * np.maximum(1.0, np.abs(E2)))
        with pytest.raises(smcb.SmcbError, match="SMCB_FP32"):
            eng.sweep(1, smcb.FP32)
