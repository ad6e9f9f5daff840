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


@pytest.mark.parametrize("N,A,T,nsweeps,kind", [(256, 0.01, 0.9, 30, "droplet"), (108, 1.1, 1.1, 60, "mixed"),
                                                (100, 0.02, 0.8, 40, "droplet"), (500, 0.02, 1.0, 6, "droplet")])
def test_fast_sweep_cache_stays_consistent(orc, N, A, T, nsweeps, kind):
    """The FAST sweep kernel keeps per-particle energy/force/neighbour caches and corrects them
    incrementally (csrc/sweep_cached.cuh).  After many sweeps in one launch - dense droplets with
    ~80 partners per particle included - the caches must equal a fresh evaluation of the final
    configuration, the neighbour counts exactly, and the running energy the recomputed one."""
    M = 3 if N != 100 else 2
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    rng = np.random.default_rng(N + nsweeps)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    nchains = 6
    if kind == "droplet":
        R0 = np.stack([config_droplet(N, L, Lz, rng, jitter=0.03, nz=4 if N <= 256 else 8) for _ in range(nchains)])
    else:
        R0 = mixed_configs(N, L, Lz, nchains, seed=N, orc=orc)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.set_rng(7, 0, 0)
        eng.debug_capture_cache(True)
        eng.sweep(nsweeps, smcb.FAST)
        ce, cf, nb = eng.debug_get_cache()
        E, na, nt = eng.chain_state()
        R = eng.get_positions()
        ev = eng.evaluate(smcb.FAST)
    assert na.sum() > 0.05 * nt.sum(), "too few accepted moves to exercise the cache updates"
    assert not np.array_equal(R, R0)
    e_tot = ev["e_lj"] + ev["e_wall"]
    f_tot = ev["f_lj"] + ev["f_wall"]
    for c in range(nchains):
        # exact neighbour counts from the oracle geometry
        X = R[c].reshape(N, 3)
        d = X[:, None, :] - X[None, :, :]
        d[:, :, 0] -= L * np.rint(d[:, :, 0] / L)
        d[:, :, 1] -= L * np.rint(d[:, :, 1] / L)
        r2 = np.einsum("ijk,ijk->ij", d, d)
        np.fill_diagonal(r2, 1e30)
        margin = np.abs(r2 - 9.0) < 1e-9
        nbo = (r2 < 9.0).sum(axis=1)
        if not margin.any():
            np.testing.assert_array_equal(nb[c], nbo)
        assert rel_err(ce[c], e_tot[c], floor=max(1.0, np.abs(e_tot[c]).max() * 1e-3)) < 1e-11
        assert rel_err(cf[c], f_tot[c], floor=max(1.0, np.abs(f_tot[c]).max() * 1e-3)) < 1e-11
        Erec = ev["U_lj"][c] + ev["U_wall"][c]
        assert abs(E[c] - Erec) <= 1e-10 * max(1.0, abs(Erec))
        # and against the oracle for one chain's per-particle forces
        if c == 0:
            fo = np.concatenate([orc.force_single(s, R[c], i) + orc.walls_force(s, R[c][3 * i:3 * i + 3], W) for i in range(N)])
            assert rel_err(cf[c], fo, floor=max(1.0, np.abs(fo).max() * 1e-3)) < 1e-11


def test_fast_sweep_multi_sweep_launch_tracks_oracle(orc):
    """several sweeps inside ONE launch (caches carried from trial to trial and sweep to sweep)
    against the oracle on the same fed numbers; chaos amplifies 1e-16 by about a decade per 10
    sweeps, so 8 sweeps must still agree to 1e-9 and take identical accept decisions"""
    N, M, T, A = 108, 3, 1.1, 1.1
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains, nsweeps = 4, 8
    rng = np.random.default_rng(31)
    R0 = mixed_configs(N, L, Lz, nchains, seed=77, orc=orc)
    streams = np.stack([make_stream(N, nsweeps, rng) for _ in range(nchains)], axis=1)
    displ, off, u = expand_streams(orc, N, A, streams)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.refresh_energy(smcb.FAST)
        E0 = eng.chain_state()[0]
        acc = eng.sweep_fed(displ, off, u, mode=smcb.FAST, want_accepted=True)
        R = eng.get_positions()
        E = eng.chain_state()[0]
    for c in range(nchains):
        Ro, Eo, tot, flags = _oracle_run(orc, s, R0[c], W, A, T, displ[:, c], off[:, c], u[:, c], E0[c])
        np.testing.assert_array_equal(acc[:, c], flags)
        assert rel_err(R[c], Ro) < 1e-9
        assert abs(E[c] - Eo) <= 1e-9 * max(1.0, abs(Eo))


@pytest.mark.parametrize("N,A", [(1024, 0.05), (600, 0.3)])
def test_fast_sweep_large_N_block_kernel(orc, N, A):
    """N > 512: the block-per-chain FAST sweep (csrc/sweep_block.cuh), teacher-forced against the oracle's
    oneParticleMoves on the same fed random numbers: same accept decisions, positions and energy change within
    1e-12; then several Philox sweeps in one launch with the running energy checked against a fresh evaluation"""
    M, T = 3, 1.1
    L, Lz = 33.0, 240.0
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains, nsweeps = 3, 3
    rng = np.random.default_rng(7 + N)
    R = mixed_configs(N, L, Lz, nchains, seed=3 * N, orc=orc)
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
                j, Eo, fl = orc.sweep(s, R[c], W, A, T, displ[0, c], off[0, c], u[0, c], Eo0, want_flags=True)
                np.testing.assert_array_equal(acc[0, c], fl)
                assert rel_err(Rg[c], R[c], floor=1.0) < 1e-12, (k, c)
                assert abs((Eg[c] - E0[c]) - (Eo - Eo0)) <= 1e-11 * max(1.0, abs(Eo - Eo0), abs(Eo0))
        assert acc.sum() > 0
        eng.set_positions(R)
        eng.set_rng(3, 0, 0)
        Et, at = eng.sweep_traced(4, smcb.FAST)
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        Erec = ev["U_lj"] + ev["U_wall"]
        assert np.all(np.abs(E - Erec) <= 1e-9 * np.maximum(1.0, np.abs(Erec)))
        np.testing.assert_array_equal(Et[-1], E)
        assert at.sum() > 0 and np.all(at <= N)


@pytest.mark.parametrize("N,A,threads", [(1500, 0.05, 1024), (1024, 0.3, 256), (4096, 0.01, 512), (650, 1.1, 512), (1100, 0.2, 128)])
def test_block_spec_sweep_equals_serial_block_sweep(orc, N, A, threads, monkeypatch):
    """the batch-speculative block sweep (csrc/sweep_block_spec.cuh: one or two trials per warp, a batch of trials evaluated
    against the same state, the valid prefix committed) against the trial-by-trial block kernel (SMCB_BLOCK_SWEEP=serial)
    on the same fed numbers over several sweeps: the accept flags of every trial are identical and the positions and
    energies agree to rounding (the two kernels add a point's pair terms in different orders), in a gas, a dense
    droplet (most batches are cut short by an accepted neighbour) and at N = 4096; free-running Philox sweeps likewise"""
    M, T = 3, 1.1
    L, Lz = 33.0, 240.0
    W = GOLDEN_W_M3.copy()
    nchains, nsweeps = 3, 3
    rng = np.random.default_rng(N)
    if N == 4096:
        X = orc.fcc_lattice(L, Lz, 16, 16, 4)
        R0 = np.stack([X + 0.02 * rng.standard_normal(3 * N) for _ in range(nchains)])
    else:
        R0 = mixed_configs(N, L, Lz, nchains, seed=5 * N, orc=orc)
    streams = np.stack([make_stream(N, nsweeps, rng) for _ in range(nchains)], axis=1)
    displ, off, u = expand_streams(orc, N, A, streams)
    out = {}
    for which in ("serial", "spec", "spec again"):
        monkeypatch.setenv("SMCB_BLOCK_SWEEP", which.split()[0])
        if which != "serial":                                  # 1024 / 128 threads: one trial per warp; 512 / 256: two
            monkeypatch.setenv("SMCB_BLOCK_SPEC_TPW", "1" if threads in (1024, 128) else "2")
            monkeypatch.setenv("SMCB_BLOCK_SWEEP_THREADS", str(threads))
        with smcb.Engine(nchains, N, M) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
            eng.set_positions(R0)
            eng.refresh_energy(smcb.FAST)
            acc = eng.sweep_fed(displ, off, u, mode=smcb.FAST, want_accepted=True)
            R, (E, na, nt) = eng.get_positions(), eng.chain_state()
            pt = eng.last_pair_counts()
            eng.set_positions(R0)
            eng.refresh_energy(smcb.FAST)
            eng.set_rng(21, 0, 0)
            eng.reset_counters()
            Et, at = eng.sweep_traced(3, smcb.FAST)
            out[which] = (acc, R, E, na, nt, Et, at, eng.get_positions(), pt)
    a, b, b2 = out["serial"], out["spec"], out["spec again"]
    for x, y in zip(b, b2):                                          # a data race in the batch protocol would show here
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    assert a[0].sum() > 0
    np.testing.assert_array_equal(a[0], b[0])
    assert rel_err(b[1], a[1], floor=1.0) < 1e-11
    assert np.all(np.abs(a[2] - b[2]) <= 1e-10 * np.maximum(1.0, np.abs(a[2])))
    np.testing.assert_array_equal(a[3], b[3]); np.testing.assert_array_equal(a[4], b[4])
    np.testing.assert_array_equal(a[6], b[6])                       # accepted moves per Philox sweep
    assert np.all(np.abs(a[5] - b[5]) <= 1e-10 * np.maximum(1.0, np.abs(a[5])))
    assert rel_err(b[7], a[7], floor=1.0) < 1e-11
    assert tuple(b[8]) == tuple(a[8])                                # nominal and in-cutoff pairs of the committed trials: the same counts


@pytest.mark.parametrize("which", ["spec", "serial"])
@pytest.mark.parametrize("N,A,nsweeps", [(600, 0.3, 4), (1024, 0.05, 3), (2000, 1.1, 1), (4096, 0.02, 1)])
def test_strict_block_sweep_bit_exact_beyond_512(orc, N, A, nsweeps, which, monkeypatch):
    """the bit-exact sweep for N > 512 (csrc/sweep_block_strict.cuh, one block per chain): free-running against the
    oracle's oneParticleMoves on the same fed numbers - accept flags, positions and running energy identical.  Both
    organisations: batch-speculative (a warp per trial, the default) and trial by trial (SMCB_BLOCK_SWEEP=serial)."""
    if which == "serial" and N == 4096:
        pytest.skip("covered by the batch-speculative kernel; the serial one takes long here")
    monkeypatch.setenv("SMCB_BLOCK_SWEEP", which)
    M, T = 3, 1.1
    L, Lz = 33.0, 240.0
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains = 2
    rng = np.random.default_rng(N)
    if N == 4096:                                           # BASELINE configs[4]: the corrected 16x16x4 fcc slab, jittered
        X = orc.fcc_lattice(L, Lz, 16, 16, 4)
        R0 = np.stack([X + 0.02 * rng.standard_normal(3 * N) for _ in range(nchains)])
    else:
        R0 = mixed_configs(N, L, Lz, nchains, seed=11 * N, orc=orc)
    streams = np.stack([make_stream(N, nsweeps, rng) for _ in range(nchains)], axis=1)
    displ, off, u = expand_streams(orc, N, A, streams)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.refresh_energy(smcb.STRICT)
        E0 = eng.chain_state()[0]
        acc = eng.sweep_fed(displ, off, u, mode=smcb.STRICT, want_accepted=True)
        R = eng.get_positions()
        E, na, nt = eng.chain_state()
    assert acc.sum() > 0
    for c in range(nchains):
        Ro, Eo, tot, flags = _oracle_run(orc, s, R0[c], W, A, T, displ[:, c], off[:, c], u[:, c], E0[c])
        np.testing.assert_array_equal(acc[:, c], flags, err_msg=f"accept flags chain {c}")
        np.testing.assert_array_equal(R[c], Ro, err_msg=f"positions chain {c}")
        assert E[c] == Eo and na[c] == tot and nt[c] == nsweeps * N


def test_block_sweep_N4096_and_bulk_mode(orc):
    """the block-per-chain sweep at the BASELINE configs[4] size (N = 4096, corrected 16x16x4 fcc lattice) and in
    bulk mode (z periodic): running energy equals a fresh evaluation, x,y stay in the box, acceptance is sane"""
    N, M = 4096, 3
    L, Lz = 33.0, 240.0
    X = orc.fcc_lattice(L, Lz, 16, 16, 4)
    rng = np.random.default_rng(1)
    R0 = np.stack([X + 0.02 * rng.standard_normal(3 * N) for _ in range(3)])
    with smcb.Engine(3, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=0.01), GOLDEN_W_M3)
        eng.set_positions(R0)
        eng.set_rng(8, 0, 0)
        eng.sweep(2, smcb.FAST)
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        R = eng.get_positions().reshape(3, N, 3)
    Erec = ev["U_lj"] + ev["U_wall"]
    assert np.all(np.abs(E - Erec) <= 1e-9 * np.maximum(1.0, np.abs(Erec)))
    assert np.all(nt == 2 * N) and np.all(na > 0.2 * nt) and np.all(na < nt)
    assert np.all(np.abs(R[:, :, :2]) <= L / 2 + 1e-12)
    # bulk: N = 864 liquid-like fcc, rho* = 0.8, 3-D minimum image, cutoff 3
    Nb = 864
    Lb = (Nb / 0.8) ** (1 / 3)
    cells = np.array([(i, j, k) for i in range(6) for j in range(6) for k in range(6)], dtype=float)
    basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    Xb = ((cells[:, None, :] + basis[None]).reshape(-1, 3) * (Lb / 6) - Lb / 2 + 0.1).reshape(-1)
    sb = make_sys(Nb, 3, Lb, Lb, periodic_z=1, wall=0)
    with smcb.Engine(2, Nb, 3) as eng:
        eng.set_params(smcb.default_params(L=Lb, Lz=Lb, T=1.0, A=0.005, flags=smcb.PERIODIC_Z))
        eng.broadcast_positions(Xb)
        eng.set_rng(9, 0, 0)
        eng.sweep(3, smcb.FAST)
        E, na, nt = eng.chain_state()
        Rb = eng.get_positions()
    for c in range(2):
        Eo = orc.energy(sb, Rb[c])
        assert abs(E[c] - Eo) <= 1e-9 * abs(Eo)
        assert np.all(np.abs(Rb[c]) <= Lb / 2 + 1e-12)          # z wrapped too
    assert na.sum() > 0
