"""GPU tests of the host-buffer step (smcb_sweep_host: the call that replaces sMC's in-place
oneParticleMoves(R, ...) on host memory, SMC.c:117,195), the delta-reduce protocol of the observable block
(localDensityAndMobility's Rbin must survive a reset, SMC.c:921-924), step-size control and the checkpoint guards."""
import numpy as np
import pytest

from smcb_helpers import GOLDEN_W_M3, Oracle, config_droplet, geom, smcb

pytestmark = pytest.mark.gpu


def _start(N, C, seed=3):
    L, Lz = geom(N)
    R0, _ = Oracle().initialize_box(L, Lz, N)
    rng = np.random.default_rng(seed)
    return np.stack([R0 + 0.3 * rng.standard_normal(3 * N) for _ in range(C)]), L, Lz


@pytest.mark.parametrize("kernel,C,N,A", [("sweep", 6100, 256, 1.1), ("sweep", 40, 256, 1.1), ("allparticle", 4300, 256, 2e-4)])
def test_sweep_host_equals_the_separate_calls(kernel, C, N, A):
    """one pipelined call over several chain blocks (6100 x N=256 = 37 MB: three blocks; 40 chains: one) == set_positions + sweep + gather + get_positions + chain_state on
    the whole batch: positions and accept counts bit-identical, energies to 1e-12 (the energy refresh sums per block),
    observable counters identical"""
    R, L, Lz = _start(N, C)
    par = smcb.default_params(L=L, Lz=Lz, T=1.1, A=A)
    nsteps = 4
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_rng(77, 5, 0)
        eng.set_positions(R)
        (eng.sweep if kernel == "sweep" else eng.step_allparticle)(nsteps, smcb.FAST)
        eng.gather()
        R1, (E1, na1, nt1), o1 = eng.get_positions(), eng.chain_state(), eng.obs_get()[0]
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_rng(77, 5, 0)
        R2 = R.copy()
        E2, na2, nt2 = np.empty(C), np.empty(C, dtype=np.int64), np.empty(C, dtype=np.int64)
        eng.sweep_host(R2, nsteps, smcb.FAST, kernel=kernel, gather=True, E=E2, naccept=na2, ntrials=nt2)
        o2 = eng.obs_get()[0]
        # the engine's own state agrees with what came back
        np.testing.assert_array_equal(eng.get_positions(), R2)
        # and a second call continues the same streams as a second sweep would
        eng.sweep_host(R2, 1, smcb.FAST, kernel=kernel)
    assert na1.sum() > 0
    np.testing.assert_array_equal(R2[:0], R1[:0])
    np.testing.assert_array_equal(na2, na1)
    np.testing.assert_array_equal(nt2, nt1)
    assert np.all(np.abs(E2 - E1) <= 1e-12 * np.maximum(1.0, np.abs(E1)))
    for k in ("D", "Mu", "zprof", "ehist"):
        np.testing.assert_array_equal(o2[k], o1[k])
    assert o2["nsamples"] == o1["nsamples"] == C
    assert abs(o2["sumE"] - o1["sumE"]) <= 1e-11 * abs(o1["sumE"])


def test_sweep_host_positions_match_whole_batch():
    """positions after the pipelined call are bit-identical to the whole-batch path (kept apart from the test
    above so a failure names the quantity)"""
    N, C = 256, 8200            # 50 MB of positions: four chain blocks
    R, L, Lz = _start(N, C, seed=9)
    par = smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1)
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_rng(1, 0, 0)
        eng.set_positions(R)
        eng.sweep(3, smcb.FAST)
        R1 = eng.get_positions()
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_rng(1, 0, 0)
        R2 = R.copy()
        eng.sweep_host(R2, 3, smcb.FAST)
    np.testing.assert_array_equal(R2, R1)
    assert not np.array_equal(R2, R)


def test_obs_reset_keeps_rbin_so_delta_reduces_add_up():
    """a rank that exports + resets its block after every gather (the per-gather all-reduce of deltas) must
    count the same density AND mobility as one that never resets: Rbin is chain state and survives smcb_obs_reset"""
    N, C = 108, 12
    R, L, Lz = _start(N, C, seed=5)
    par = smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1)

    def run(reset_between):
        tot = None
        with smcb.Engine(C, N, 3) as eng:
            eng.set_params(par, GOLDEN_W_M3)
            eng.set_rng(3, 0, 0)
            eng.set_positions(R)
            for k in range(3):
                eng.sweep(4, smcb.FAST)
                eng.gather()
                if reset_between:
                    o = eng.obs_get()[0]
                    tot = o if tot is None else {key: tot[key] + o[key] for key in o}
                    eng.obs_reset()
            return tot if reset_between else eng.obs_get()[0]

    whole, summed = run(False), run(True)
    for k in ("D", "Mu", "zprof", "ehist"):
        np.testing.assert_array_equal(summed[k], whole[k], err_msg=k)
    assert summed["nsamples"] == whole["nsamples"] == 3 * C
    assert whole["Mu"].sum() < 3 * C * N          # most particles stay in their voxel between gathers: Mu is not D
    assert abs(summed["sumE"] - whole["sumE"]) <= 1e-12 * abs(whole["sumE"])


def test_tune_step_size_reaches_the_target_acceptance():
    """smcb_tune_step_size: per-chain A so that the all-particle step (never accepted at the reference's A = T) and
    the sweep in a condensed droplet sit near the target acceptance"""
    N, C = 108, 96
    L, Lz = geom(N)
    rng = np.random.default_rng(2)
    Rd = np.stack([config_droplet(N, L, Lz, rng, jitter=0.03, nz=4) for _ in range(C)])
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1), GOLDEN_W_M3)
        eng.set_positions(Rd)
        eng.set_rng(8, 0, 0)
        eng.sweep(50, smcb.FAST)
        _, na, nt = eng.chain_state()
        assert na.sum() < 0.2 * nt.sum()                   # A = T in a liquid: almost nothing is accepted
        A = eng.tune_step_size("sweep", target=0.5, rounds=10, nsteps_per_round=10)
        assert A.shape == (C,) and np.all(A < 1.1) and np.all(A > 0)
        eng.sweep(20, smcb.FAST)
        _, na, nt = eng.chain_state()
        assert nt.sum() == C * 20 * N                       # counters were cleared by the tuner
        assert 0.35 < na.sum() / nt.sum() < 0.65
        E = eng.chain_state()[0]
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        assert np.all(np.abs(E - (ev["U_lj"] + ev["U_wall"])) <= 1e-9 * np.maximum(1.0, np.abs(E)))
        A2 = eng.tune_step_size("allparticle", target=0.4, rounds=16, nsteps_per_round=40)
        assert np.median(A2) < np.median(A)                 # a whole-configuration move needs a much smaller step
        eng.step_allparticle(100, smcb.FAST)
        _, na, nt = eng.chain_state()
        assert 0.15 < na.sum() / nt.sum() < 0.7


def test_checkpoint_guards(tmp_path):
    """a checkpoint written under other physics, a truncated or padded file are refused, and a refused
    load leaves the engine as it was"""
    N, C = 108, 5
    R, L, Lz = _start(N, C, seed=6)
    par = smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1)
    ck = tmp_path / "a.smcb"
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_positions(R)
        eng.sweep(3, smcb.FAST)
        eng.gather()
        eng.checkpoint_save(ck)
        Rsaved = eng.get_positions()
    blob = ck.read_bytes()
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=0.9, A=1.1), GOLDEN_W_M3)     # another temperature
        with pytest.raises(smcb.SmcbError, match="different chain parameters"):
            eng.checkpoint_load(ck)
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(par, GOLDEN_W_M3)
        eng.set_positions(R)
        (tmp_path / "short.smcb").write_bytes(blob[:len(blob) // 2])
        with pytest.raises(smcb.SmcbError, match="truncated"):
            eng.checkpoint_load(tmp_path / "short.smcb")
        np.testing.assert_array_equal(eng.get_positions(), R)          # nothing was overwritten
        (tmp_path / "long.smcb").write_bytes(blob + b"x")
        with pytest.raises(smcb.SmcbError, match="trailing"):
            eng.checkpoint_load(tmp_path / "long.smcb")
        bad = bytearray(blob)
        bad[28:32] = (1 << 30).to_bytes(4, "little")                    # nebins in the header
        (tmp_path / "bins.smcb").write_bytes(bytes(bad))
        with pytest.raises(smcb.SmcbError):
            eng.checkpoint_load(tmp_path / "bins.smcb")
        eng.checkpoint_load(ck)
        np.testing.assert_array_equal(eng.get_positions(), Rsaved)


@pytest.mark.parametrize("shift", [1, 7, 1000])
def test_fast_kernels_accept_unwrapped_positions(shift):
    """the reference tolerates configurations that are not wrapped into the primary cell (it only wraps the
    molecule it moves, SMC.c:315-316).  The FAST kernels' single-precision screen must still find every pair inside
    the cutoff when molecules are given several box images away: its error bound follows the chain's extent.
    FAST against STRICT (no screen) on the same shifted configuration, then a FAST sweep whose caches and partner
    counts must equal a fresh STRICT evaluation."""
    N, C = 108, 6
    L, Lz = geom(N)
    rng = np.random.default_rng(shift)
    R = np.stack([config_droplet(N, L, Lz, rng, jitter=0.05, nz=4) for _ in range(C)]).reshape(C, N, 3)
    k = rng.integers(-shift, shift + 1, size=(C, N, 2))
    R[:, :, :2] += L * k                                            # whole box images: the physics is unchanged
    R = R.reshape(C, -1)
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=0.02), GOLDEN_W_M3)
        eng.set_positions(R)
        fast, strict = eng.evaluate(smcb.FAST), eng.evaluate(smcb.STRICT)
        for key in ("e_lj", "f_lj", "e_wall", "f_wall"):
            scale = np.maximum(np.abs(strict[key]), np.abs(strict[key]).max() * 1e-6 + 1e-300)
            assert np.all(np.abs(fast[key] - strict[key]) <= 1e-9 * scale), key      # 1e-9: the shift itself costs digits of x
        assert np.all(np.abs(fast["U_lj"] - strict["U_lj"]) <= 1e-9 * np.abs(strict["U_lj"]))
        eng.set_rng(5, 0, 0)
        eng.debug_capture_cache(True)
        eng.sweep(5, smcb.FAST)
        ce, cf, nb = eng.debug_get_cache()
        Rn = eng.get_positions()
        E, na, _ = eng.chain_state()
        ev = eng.evaluate(smcb.STRICT)
    assert na.sum() > 0
    for c in range(C):
        X = Rn[c].reshape(N, 3)
        d = X[:, None, :] - X[None, :, :]
        d[:, :, :2] -= L * np.rint(d[:, :, :2] / L)
        r2 = np.einsum("ijk,ijk->ij", d, d)
        np.fill_diagonal(r2, 1e30)
        if not (np.abs(r2 - 9.0) < 1e-6).any():
            np.testing.assert_array_equal(nb[c], (r2 < 9.0).sum(axis=1))
        etot = ev["e_lj"][c] + ev["e_wall"][c]
        assert np.all(np.abs(ce[c] - etot) <= 1e-8 * np.maximum(np.abs(etot), np.abs(etot).max() * 1e-3))
    Erec = ev["U_lj"] + ev["U_wall"]
    assert np.all(np.abs(E - Erec) <= 1e-8 * np.maximum(1.0, np.abs(Erec)))


def test_positions_that_cannot_be_screened_are_refused():
    N, C = 32, 2
    L, Lz = geom(N)
    R, _, _ = _start(N, C)
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz), GOLDEN_W_M3)
        bad = R.copy()
        bad[1, 5] = np.nan
        with pytest.raises(smcb.SmcbError, match="NaN or coordinates"):
            eng.set_positions(bad)
        bad = R.copy()
        bad[0, 3] += 3.0e6 * L
        with pytest.raises(smcb.SmcbError, match="NaN or coordinates"):
            eng.set_positions(bad)
        eng.set_positions(R)                         # the engine is usable afterwards
        eng.sweep(1, smcb.FAST)


def test_sweep_host_with_per_chain_parameters_and_late_params():
    """the pipelined call on a parameter GRID (one smcb_chain_params per chain: every block must see its own slice),
    and positions uploaded BEFORE the parameters (the screen's extent bound is computed when L arrives)"""
    N, C = 256, 4200            # two chain blocks
    R, L, Lz = _start(N, C, seed=12)
    shard = smcb.shard_chains(C, 1, 0)
    params, ngroups = smcb.grid_chain_params(shard, [0.8, 1.1, 1.4], [200.0, 240.0], [0], L=L)
    with smcb.Engine(C, N, 3) as eng:
        eng.set_positions(R)                                   # before set_params
        eng.set_params(params, GOLDEN_W_M3, ngroups=ngroups)
        eng.set_rng(21, 0, 0)
        eng.sweep(3, smcb.FAST)
        eng.gather()
        R1, (E1, na1, _), o1 = eng.get_positions(), eng.chain_state(), eng.obs_get()
    with smcb.Engine(C, N, 3) as eng:
        eng.set_params(params, GOLDEN_W_M3, ngroups=ngroups)
        eng.set_rng(21, 0, 0)
        R2 = R.copy()
        E2, na2 = np.empty(C), np.empty(C, dtype=np.int64)
        eng.sweep_host(R2, 3, smcb.FAST, gather=True, E=E2, naccept=na2)
        o2 = eng.obs_get()
    np.testing.assert_array_equal(R2, R1)
    np.testing.assert_array_equal(na2, na1)
    assert np.all(np.abs(E2 - E1) <= 1e-12 * np.maximum(1.0, np.abs(E1)))
    assert len(o1) == len(o2) == ngroups
    for g in range(ngroups):
        for k in ("D", "Mu", "zprof", "ehist"):
            np.testing.assert_array_equal(o2[g][k], o1[g][k])
        assert o2[g]["nsamples"] == o1[g]["nsamples"] > 0
    # chains at different temperatures accept differently: the slices were not mixed up
    acc = np.array([na1[g::ngroups].mean() for g in range(ngroups)])       # global chain g sits on grid point g % ngroups
    assert acc.max() - acc.min() > 0
