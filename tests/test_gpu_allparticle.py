"""GPU parity, the all-particle Smart-MC step (north-star kernel B) and the Philox streams.

There is no working reference for this step (markovProbability is dead code, SMC.c:354-402); it
is pinned against the CPU restatement orc_allparticle_step, which is assembled from the
reference's own forceSingle/wallsForce/energy/wallsEnergy (already pinned bit-exactly), with the
same host-fed noise; and its sampling is pinned statistically in test_gpu_observables.py."""
import numpy as np
import pytest

from smcb_helpers import GOLDEN_W_M3, Oracle, geom, make_sys, mixed_configs, rel_err, smcb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("mode", [smcb.STRICT, smcb.FAST])
@pytest.mark.parametrize("N,A,nsteps", [(108, 2e-4, 30), (256, 1e-4, 12), (32, 1e-3, 60), (101, 2e-4, 20), (33, 1e-3, 40),
                                        (500, 5e-5, 8)])
def test_allparticle_fed_matches_oracle(orc, mode, N, A, nsteps):
    M, T = 3, 1.1
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains = 4
    rng = np.random.default_rng(N + nsteps)
    R0 = mixed_configs(N, L, Lz, nchains, seed=17 * N, orc=orc)
    xi = rng.standard_normal((nsteps, nchains, 3 * N)) * np.sqrt(2 * A)
    u = rng.random((nsteps, nchains))
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        lnap, acc = eng.step_allparticle_fed(xi, u, mode=mode)
        R = eng.get_positions()
        E, na, nt = eng.chain_state()
    assert acc.sum() > 0, "nothing accepted: test would be vacuous"
    for c in range(nchains):
        Ro = R0[c].copy()
        F, Ulj, Uw, _ = orc.total(s, Ro, W)
        U = Ulj + Uw
        nacc = 0
        for k in range(nsteps):
            ok, U, ln = orc.allparticle_step(s, Ro, F, U, W, A, T, xi[k, c], u[k, c])
            assert abs(lnap[k, c] - ln) <= 1e-9 * max(1.0, abs(ln)), (k, c, lnap[k, c], ln)
            assert bool(acc[k, c]) == bool(ok), (k, c)
            nacc += ok
        assert rel_err(R[c], Ro) < 1e-11
        assert abs(E[c] - U) <= 1e-10 * max(1.0, abs(U))
        assert na[c] == nacc and nt[c] == nsteps


def test_philox_stream_matches_oracle(orc):
    """Production mode draws from Philox4x32-10; the CPU restatement of the stream
    (orc_rng_particle / orc_rng_step_scalars) fed to the oracle sweep must give the same sweep."""
    N, M, T, A = 64, 3, 1.1, 0.4
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains, nsweeps, seed, chain0, step0 = 3, 5, 0x1234ABCD5678EF01, 1000, 77
    R0 = mixed_configs(N, L, Lz, nchains, seed=4, orc=orc)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.set_rng(seed, chain0, step0)
        eng.refresh_energy(smcb.STRICT)
        E0 = eng.chain_state()[0]
        eng.sweep(nsweeps, smcb.STRICT)
        R = eng.get_positions()
        E, na, _ = eng.chain_state()
    sigma = np.sqrt(2.0 * A)
    for c in range(nchains):
        Ro, Eo, tot = R0[c].copy(), E0[c], 0
        for k in range(nsweeps):
            off, _ = orc.rng_step_scalars(seed, chain0 + c, step0 + k)
            displ = np.empty(3 * N)
            ubyp = np.empty(N)
            for n in range(N):
                g, un = orc.rng_particle(seed, chain0 + c, step0 + k, n)
                displ[3 * n:3 * n + 3] = sigma * g
                ubyp[n] = un
            order = (np.arange(N) + off) % N
            j, Eo = orc.sweep(s, Ro, W, A, T, displ, off, ubyp[order].copy(), Eo)
            tot += j
        assert na[c] == tot
        assert rel_err(R[c], Ro) < 1e-9          # device log/sincospi differ from libm by ulps
        assert abs(E[c] - Eo) <= 1e-9 * max(1.0, abs(Eo))


def test_allparticle_philox_energy_bookkeeping(orc):
    """free-running kernel B: the cached energy equals a fresh evaluation afterwards, forces cached
    across calls are reused (second call does not refresh)"""
    N, M, T, A = 256, 3, 1.1, 1e-4
    L, Lz = geom(N)
    W = GOLDEN_W_M3.copy()
    nchains = 16
    R0 = mixed_configs(N, L, Lz, nchains, seed=21, orc=orc)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.set_rng(5, 0, 0)
        eng.step_allparticle(40, smcb.FAST)
        eng.step_allparticle(40, smcb.FAST)
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        tot, cut = eng.last_pair_counts()
    assert np.all(nt == 80) and na.sum() > 0
    Erec = ev["U_lj"] + ev["U_wall"]
    assert np.all(np.abs(E - Erec) <= 1e-10 * np.maximum(1.0, np.abs(Erec)))


@pytest.mark.parametrize("cluster", [1, 2, 4])
def test_allparticle_large_N_cluster_matches_oracle(orc, cluster, monkeypatch):
    """N = 1024 with the chain split over a thread-block cluster (DSMEM reduction of the MH sums):
    every cluster size must reproduce the oracle's step (ln ap within 1e-9, same decisions)"""
    from oracle_bindings import config_droplet, config_gas
    monkeypatch.setenv("SMCB_CLUSTER", str(cluster))
    N, M, T, A, nsteps = 1024, 3, 1.1, 2e-5, 5
    L, Lz = 33.0, 240.0
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    rng = np.random.default_rng(99)
    R0 = np.stack([config_droplet(N, L, Lz, rng, jitter=0.04, nz=8), config_gas(N, L, Lz, rng)])
    xi = rng.standard_normal((nsteps, 2, 3 * N)) * np.sqrt(2 * A)
    u = rng.random((nsteps, 2))
    with smcb.Engine(2, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        lnap, acc = eng.step_allparticle_fed(xi, u, mode=smcb.FAST)
        R = eng.get_positions()
        E, na, nt = eng.chain_state()
    for c in range(2):
        Ro = R0[c].copy()
        F, Ulj, Uw, _ = orc.total(s, Ro, W)
        U = Ulj + Uw
        for k in range(nsteps):
            ok, U, ln = orc.allparticle_step(s, Ro, F, U, W, A, T, xi[k, c], u[k, c])
            assert abs(lnap[k, c] - ln) <= 1e-9 * max(1.0, abs(ln)), (k, c, lnap[k, c], ln)
            assert bool(acc[k, c]) == bool(ok), (k, c)
        assert rel_err(R[c], Ro) < 1e-11
        assert abs(E[c] - U) <= 1e-10 * max(1.0, abs(U))


def test_allparticle_N4096_config5_invariants(orc):
    """BASELINE configs[4] shape (N = 4096 with wall; clusters picked automatically for a small batch):
    the energy carried by the kernel equals a fresh evaluation, and a second call reuses the forces"""
    N, M, T, A = 4096, 3, 1.1, 2e-6
    L, Lz = 33.0, 240.0
    W = GOLDEN_W_M3.copy()
    X = orc.fcc_lattice(L, Lz, 16, 16, 4)                   # the corrected N=4096 lattice
    assert X.size == 3 * N
    rng = np.random.default_rng(3)
    R0 = np.stack([X + 0.02 * rng.standard_normal(3 * N) for _ in range(4)])
    with smcb.Engine(4, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), W)
        eng.set_positions(R0)
        eng.set_rng(11, 0, 0)
        eng.step_allparticle(6, smcb.FAST)
        eng.step_allparticle(6, smcb.FAST)
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        tot, cut = eng.last_pair_counts()
    assert np.all(nt == 12) and na.sum() > 0
    assert tot == 4 * 6 * N * (N - 1) and cut > 0
    Erec = ev["U_lj"] + ev["U_wall"]
    assert np.all(np.abs(E - Erec) <= 1e-10 * np.maximum(1.0, np.abs(Erec)))
