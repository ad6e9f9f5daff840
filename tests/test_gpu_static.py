"""GPU parity, static routines: smcb_evaluate through the C ABI against
the oracle on the same configurations.
  STRICT: per-particle energySingle / forceSingle / wallsEnergySingle / wallsForce BIT-IDENTICAL
          (SMC.c:557-618, 729-813); chain totals within 1e-12 relative (different summation tree).
  FAST  : everything within 1e-12 relative (north_star tolerance, fp64)."""
import numpy as np
import pytest

from smcb_helpers import GOLDEN_W_M3, Oracle, geom, make_sys, mixed_configs, random_walls, rel_err, smcb

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _oracle_eval(orc, s, R, W):
    N = s.N
    e_lj = np.array([orc.energy_single(s, R, i) for i in range(N)])
    f_lj = np.concatenate([orc.force_single(s, R, i) for i in range(N)])
    e_w = np.array([orc.walls_energy_single(s, R[3 * i:3 * i + 3], W) for i in range(N)])
    f_w = np.concatenate([orc.walls_force(s, R[3 * i:3 * i + 3], W) for i in range(N)])
    vol3 = 3 * s.L * s.L * s.Lz
    return dict(e_lj=e_lj, f_lj=f_lj, e_wall=e_w, f_wall=f_w, U_lj=orc.energy(s, R), U_wall=orc.walls_energy(s, R, W),
                vir_lj=-orc.pressure(s, R) * vol3, vir_wall_ref=-orc.walls_pressure(s, R, W) * vol3)


@pytest.mark.parametrize("N,M,nchains", [(32, 3, 6), (108, 3, 8), (256, 3, 8), (108, 4, 4), (500, 3, 4), (100, 2, 4)])
def test_evaluate_matches_oracle(orc, N, M, nchains):
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    rng = np.random.default_rng(N * 31 + M)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    R = mixed_configs(N, L, Lz, nchains, seed=N + M, orc=orc)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz), W)
        eng.set_positions(R)
        strict = eng.evaluate(smcb.STRICT)
        fast = eng.evaluate(smcb.FAST)
    for c in range(nchains):
        ref = _oracle_eval(orc, s, R[c], W)
        for k in ("e_lj", "f_lj", "e_wall", "f_wall"):
            np.testing.assert_array_equal(strict[k][c], ref[k], err_msg=f"STRICT {k} chain {c}")
            scale = max(1.0, float(np.max(np.abs(ref[k]))) * 1e-3)
            assert rel_err(fast[k][c], ref[k], floor=scale) < TOL, (k, c)
        for k in ("U_lj", "U_wall", "vir_lj", "vir_wall_ref"):
            for res in (strict, fast):
                assert abs(res[k][c] - ref[k]) <= TOL * max(1.0, abs(ref[k])), (k, c, res[k][c], ref[k])


def test_evaluate_bulk_periodic(orc):
    """config 1: bulk N=108, rho*=0.5 (L=6), 3-D minimum image, cutoff L/2, no wall"""
    N, L = 108, (108 / 0.5) ** (1.0 / 3.0)
    s = make_sys(N, 3, L, L, rc2=L * L / 4, periodic_z=1, wall=0)
    rng = np.random.default_rng(7)
    a = L / 3
    cells = np.array([(i, j, k) for i in range(3) for j in range(3) for k in range(3)], dtype=float)
    basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    X0 = ((cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + a / 4)
    X0 -= L * np.rint(X0 / L)
    nch = 4
    R = np.stack([(X0 + (rng.random(X0.shape) - 0.5) * 0.1 * (c + 1)).reshape(-1) for c in range(nch)])
    with smcb.Engine(nch, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=L, T=1.0, rc2=L * L / 4, flags=smcb.PERIODIC_Z))
        eng.set_positions(R)
        strict = eng.evaluate(smcb.STRICT)
        fast = eng.evaluate(smcb.FAST)
    for c in range(nch):
        e = np.array([orc.energy_single(s, R[c], i) for i in range(N)])
        f = np.concatenate([orc.force_single(s, R[c], i) for i in range(N)])
        np.testing.assert_array_equal(strict["e_lj"][c], e)
        np.testing.assert_array_equal(strict["f_lj"][c], f)
        assert rel_err(fast["f_lj"][c], f, floor=max(1.0, np.max(np.abs(f)) * 1e-3)) < TOL
        U = orc.energy(s, R[c])
        assert abs(fast["U_lj"][c] - U) <= TOL * abs(U)
        assert abs(strict["U_lj"][c] - U) <= TOL * abs(U)
        assert strict["U_wall"][c] == 0.0 and np.all(strict["f_wall"][c] == 0.0)


def test_wall_clamp_and_per_chain_params(orc):
    """particles on/behind the walls (dz = +-1e-4 clamp, ~1e40 terms, SMC.c:738-739) and
    per-chain L, Lz, wall table, cutoff"""
    N, M = 32, 3
    rng = np.random.default_rng(11)
    W = np.concatenate([GOLDEN_W_M3, random_walls(M, rng)])
    geoms = [(33.0, 200.0, 9.0, 0), (20.0, 120.0, 9.0, 1), (33.0, 60.0, 6.25, 1)]
    params, Rs = [], []
    for L, Lz, rc2, w in geoms:
        params.append(smcb.default_params(L=L, Lz=Lz, rc2=rc2, wall=w))
        R = np.empty(3 * N)
        R[0::3] = (rng.random(N) - 0.5) * L
        R[1::3] = (rng.random(N) - 0.5) * L
        R[2::3] = (rng.random(N) - 0.5) * Lz
        R[2], R[5], R[8], R[11] = -Lz / 2, Lz / 2, -Lz / 2 - 1.5, Lz / 2 + 0.3
        Rs.append(R)
    R = np.stack(Rs)
    with smcb.Engine(3, N, M) as eng:
        eng.set_params(params, W)
        eng.set_positions(R)
        strict = eng.evaluate(smcb.STRICT)
        fast = eng.evaluate(smcb.FAST)
    for c, (L, Lz, rc2, w) in enumerate(geoms):
        s = make_sys(N, M, L, Lz, rc2=rc2)
        Wc = W[w * 18:(w + 1) * 18].copy()
        ref = _oracle_eval(orc, s, R[c], Wc)
        for k in ("e_lj", "f_lj", "e_wall", "f_wall"):
            np.testing.assert_array_equal(strict[k][c], ref[k], err_msg=f"{k} chain {c}")
            assert np.all(np.abs(fast[k][c] - ref[k]) <= TOL * np.maximum(np.abs(ref[k]), 1.0)), (k, c)


def test_no_cpu_fallback_error_paths():
    with pytest.raises(smcb.SmcbError):
        smcb.Engine(0, 10)
    with smcb.Engine(2, 32) as eng:
        with pytest.raises(smcb.SmcbError):
            eng.sweep(1)                       # params / positions not set
        eng.set_params(smcb.default_params(), GOLDEN_W_M3)
        with pytest.raises(smcb.SmcbError):
            eng.evaluate()
        info = eng.device_info()
        assert info["cc"][0] >= 10 and info["sm_count"] > 0


def test_wall_virial_as_intended(orc):
    """the corrected wall virial (no reference code - wallsPressure has three defects): the CUDA
    value against the CPU restatement and against r dV/dr differentiated numerically term by term; the gathered
    pressure moment switches between the reference's arithmetic and the corrected one"""
    N, M = 108, 3
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    R = mixed_configs(N, L, Lz, 4, seed=9, orc=orc)
    with smcb.Engine(4, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz), W)
        eng.set_positions(R)
        for mode in (smcb.STRICT, smcb.FAST):
            ev = eng.evaluate(mode, per_particle=False)
            vw = eng.wall_virial()
            for c in range(4):
                ref = orc.walls_virial_intended(s, R[c], W)
                assert abs(vw[c] - ref) <= TOL * max(1.0, abs(ref)), (mode, c, vw[c], ref)
        # independent check on one chain: sum of r dV/dr with dV/dr by central differences of each 12-6 term
        c, tot, h = 2, 0.0, 1e-6
        v12_6 = lambda a, b, r: 4.0 * (a / r ** 12 - b / r ** 6)
        for n in range(N):
            x, y, z = R[c][3 * n:3 * n + 3]
            dz = z + Lz / 2
            dz -= Lz * np.rint(dz / Lz)
            tot += abs(dz) * (v12_6(s.a0, s.b0, abs(dz) + h) - v12_6(s.a0, s.b0, abs(dz) - h)) / (2 * h)
            for i in range(M):
                for j in range(M):
                    dx = x - i * L / M; dx -= L * np.rint(dx / L)
                    dy = y - j * L / M; dy -= L * np.rint(dy / L)
                    r = np.sqrt(dx * dx + dy * dy + dz * dz)
                    if r * r < 9.0:
                        a, b = W[2 * (i * M + j)], W[2 * (i * M + j) + 1]
                        tot += r * (v12_6(a, b, r + h) - v12_6(a, b, r - h)) / (2 * h)
        assert abs(vw[c] - tot) <= 1e-6 * max(1.0, abs(tot))
        # pressure moment of the gather: reference arithmetic by default, corrected on request
        vol3 = 3 * L * L * Lz
        eng.gather()
        p_ref = eng.obs_get()[0]["sumP"]
        eng.obs_reset()
        eng.obs_set_wall_virial(True)
        eng.gather()
        p_int = eng.obs_get()[0]["sumP"]
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        assert abs(p_ref - np.sum(-(ev["vir_lj"] + ev["vir_wall_ref"]) / vol3)) <= 1e-10 * max(1e-6, abs(p_ref))
        assert abs(p_int - np.sum(-(ev["vir_lj"] + eng.wall_virial()) / vol3)) <= 1e-10 * max(1e-6, abs(p_int))
