"""GPU parity, the sweep (row a1: oneParticleMoves, SMC.c:278-351) through smcb_sweep_fed.

STRICT mode fed the random numbers the reference draws must reproduce the reference trajectory
BIT FOR BIT, free-running, for >= 1e4 sweeps (north_star: "step-for-step for at least 1e4 steps").
FAST mode is checked teacher-forced (re-synchronised every sweep): 1e-12 relative per sweep."""
import numpy as np
import pytest

from smcb_helpers import (GOLDEN_W_M3, Oracle, config_droplet, expand_streams, geom, make_stream, make_sys,
                          mixed_configs, random_walls, rel_err, smcb)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _oracle_run(orc, s, R0, W, A, T, displ, off, u, E0):
    """displ [S,3N] off [S] u [S,N] for ONE chain; returns R, E, total accepted, flags [S,N]"""
    R = R0.copy()
    E, tot = E0, 0
    flags = np.zeros((displ.shape[0], s.N), dtype=np.uint8)
    for k in range(displ.shape[0]):
        j, E, fl = orc.sweep(s, R, W, A, T, displ[k], off[k], u[k], E, want_flags=True)
        flags[k] = fl
        tot += j
    return R, E, tot, flags


@pytest.mark.parametrize("N,M,T,A,nsweeps,nchains", [(108, 3, 1.1, 1.1, 300, 4), (32, 3, 0.9, 0.9, 500, 3),
                                                     (256, 3, 1.1, 0.02, 40, 4), (108, 4, 0.8, 0.01, 100, 3),
                                                     (100, 2, 1.1, 0.3, 100, 3), (500, 3, 1.1, 0.05, 6, 2)])
def test_strict_sweeps_bit_exact(orc, N, M, T, A, nsweeps, nchains):
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    rng = np.random.default_rng(N + M + nsweeps)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    R0 = mixed_configs(N, L, Lz, nchains, seed=3 * N + M, orc=orc)
    streams = np.stack([make_stream(N, nsweeps, rng) for _ in range(nchains)], axis=1)     # [S,C,4N+1]
    displ, off, u = expand_streams(orc, N, A, streams)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.refresh_energy(smcb.STRICT)
        E0, _, _ = eng.chain_state()
        acc = eng.sweep_fed(displ, off, u, mode=smcb.STRICT, want_accepted=True)
        R = eng.get_positions()
        E, na, nt = eng.chain_state()
    for c in range(nchains):
        Eo0 = orc.energy(s, R0[c]) + orc.walls_energy(s, R0[c], W)
        assert abs(E0[c] - Eo0) <= 1e-12 * max(1.0, abs(Eo0))
        Ro, Eo, tot, flags = _oracle_run(orc, s, R0[c], W, A, T, displ[:, c], off[:, c], u[:, c], E0[c])
        np.testing.assert_array_equal(acc[:, c], flags, err_msg=f"accept flags chain {c}")
        np.testing.assert_array_equal(R[c], Ro, err_msg=f"positions chain {c}")
        assert E[c] == Eo and na[c] == tot and nt[c] == nsweeps * N


def test_strict_trajectory_1e4_sweeps(orc):
    """north_star: trajectories fed the same random numbers agree step-for-step for >= 1e4 steps.
    N=108, main.c geometry, T=A=1.1, from the fcc lattice; compared at every 1000th sweep."""
    N, M, T, A = 108, 3, 1.1, 1.1
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    R0, sites = orc.initialize_box(L, Lz, N)
    assert sites == N
    rng = np.random.default_rng(2024)
    total, chunk = 10000, 1000
    Ro = R0.copy()
    with smcb.Engine(1, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0[None, :])
        eng.refresh_energy(smcb.STRICT)
        Eo = eng.chain_state()[0][0]
        tot = 0
        for blk in range(total // chunk):
            streams = make_stream(N, chunk, rng)[:, None, :]
            displ, off, u = expand_streams(orc, N, A, streams)
            eng.sweep_fed(displ, off, u, mode=smcb.STRICT)
            for k in range(chunk):
                j, Eo = orc.sweep(s, Ro, W, A, T, displ[k, 0], off[k, 0], u[k, 0], Eo)
                tot += j
            R = eng.get_positions()[0]
            E, na, _ = eng.chain_state()
            np.testing.assert_array_equal(R, Ro, err_msg=f"diverged before sweep {(blk + 1) * chunk}")
            assert E[0] == Eo and na[0] == tot
    assert 0.5 < tot / (total * N) < 1.0


@pytest.mark.parametrize("N,A,start", [(108, 1.1, "mixed"), (256, 0.02, "mixed"), (64, 0.5, "mixed")])
def test_fast_sweep_teacher_forced(orc, N, A, start):
    """FAST arithmetic, re-synchronised to the oracle state before every sweep: same accept
    decisions, positions and energy change within 1e-12 relative."""
    M, T = 3, 1.1
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains, nsweeps = 4, 25
    rng = np.random.default_rng(99 + N)
    R = mixed_configs(N, L, Lz, nchains, seed=5 * N, orc=orc)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        for k in range(nsweeps):
            streams = np.stack([make_stream(N, 1, rng) for _ in range(nchains)], axis=1)
            displ, off, u = expand_streams(orc, N, A, streams)
            eng.set_positions(R)
            eng.refresh_energy(smcb.FAST)
            E0 = eng.chain_state()[0]
            acc = eng.sweep_fed(displ, off, u, mode=smcb.FAST, want_accepted=True)
            Rg = eng.get_positions()
            Eg = eng.chain_state()[0]
            for c in range(nchains):
                Eo0 = orc.energy(s, R[c]) + orc.walls_energy(s, R[c], W)
                assert abs(E0[c] - Eo0) <= 1e-12 * max(1.0, abs(Eo0))
                j, Eo, fl = orc.sweep(s, R[c], W, A, T, displ[0, c], off[0, c], u[0, c], Eo0, want_flags=True)
                np.testing.assert_array_equal(acc[0, c], fl)
                assert rel_err(Rg[c], R[c], floor=1.0) < 1e-12, (k, c)
                assert abs((Eg[c] - E0[c]) - (Eo - Eo0)) <= 1e-12 * max(1.0, abs(Eo - Eo0), abs(Eo0))


def test_thermalisation_scale_and_bulk_mode(orc):
    """sMC thermalises with A*2 (SMC.c:110); the bulk switch wraps z as well"""
    N, M, T, A = 32, 3, 1.0, 0.05
    L = (N / 0.5) ** (1 / 3.0)
    s = make_sys(N, M, L, L, rc2=L * L / 4, periodic_z=1, wall=0)
    rng = np.random.default_rng(1)
    # jittered simple cubic start inside the periodic cube
    g = np.array([(i, j, k) for i in range(4) for j in range(4) for k in range(2)], dtype=float)
    R0 = ((g + 0.5) * np.array([L / 4, L / 4, L / 2]) - L / 2 + (rng.random(g.shape) - 0.5) * 0.1).reshape(1, -1)
    nsweeps = 50
    streams = make_stream(N, nsweeps, rng)[:, None, :]
    displ, off, u = expand_streams(orc, N, 2 * A, streams)
    with smcb.Engine(1, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=L, T=T, A=A, rc2=L * L / 4, flags=smcb.PERIODIC_Z))
        eng.set_step_scale(2.0)
        eng.set_positions(R0)
        eng.refresh_energy(smcb.STRICT)
        E0 = eng.chain_state()[0][0]
        eng.sweep_fed(displ, off, u, mode=smcb.STRICT)
        R = eng.get_positions()[0]
        E = eng.chain_state()[0][0]
    Ro = R0[0].copy()
    Eo = E0
    for k in range(nsweeps):
        _, Eo = orc.sweep(s, Ro, None if False else np.zeros(18), 2 * A, T, displ[k, 0], off[k, 0], u[k, 0], Eo)
    np.testing.assert_array_equal(R, Ro)
    assert E == Eo
    assert np.all(np.abs(R) <= L / 2 + 1e-12)
