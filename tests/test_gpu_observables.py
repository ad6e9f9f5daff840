"""GPU parity, observables (row a13 localDensityAndMobility, SMC.c:912-927, and what sMC harvests at
a gather, SMC.c:137-141) plus size-independent properties at BASELINE's full batch size."""
import numpy as np
import pytest

from smcb_helpers import GOLDEN_W_M3, Oracle, config_gas, geom, make_sys, mixed_configs, smcb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_gather_matches_local_density(orc):
    N, M = 108, 3
    L, Lz = geom(N)
    s = make_sys(N, M, L, Lz)
    W = GOLDEN_W_M3.copy()
    nchains, ngroups, ngather = 6, 2, 4
    rng = np.random.default_rng(3)
    params = [smcb.default_params(L=L, Lz=Lz, group=c % ngroups) for c in range(nchains)]
    D = np.zeros((ngroups, 33 ** 3), dtype=np.uint64)
    Mu = np.zeros_like(D)
    Rbin = np.zeros((nchains, N), dtype=np.int32)
    sumE = np.zeros(ngroups)
    sumP = np.zeros(ngroups)
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(params, W, ngroups=ngroups)
        eng.obs_configure(nebins=32, e_lo=-4.0, e_hi=1.0)
        for _ in range(ngather):
            R = np.stack([config_gas(N, L, Lz, rng, zfrac=0.499) for _ in range(nchains)])
            eng.set_positions(R)
            eng.gather()
            for c in range(nchains):
                g = c % ngroups
                orc.local_density(s, R[c], D[g], Rbin[c], Mu[g])
                sumE[g] += orc.energy(s, R[c]) + orc.walls_energy(s, R[c], W)
                sumP[g] += orc.pressure(s, R[c]) + orc.walls_pressure(s, R[c], W)
        obs = eng.obs_get()
        rb = eng.rbin()
    np.testing.assert_array_equal(rb, Rbin)
    for g in range(ngroups):
        np.testing.assert_array_equal(obs[g]["D"].reshape(-1), D[g])
        np.testing.assert_array_equal(obs[g]["Mu"].reshape(-1), Mu[g])
        np.testing.assert_array_equal(obs[g]["zprof"], D[g].reshape(33, 33, 33).sum(axis=(0, 1)))
        assert obs[g]["nsamples"] == ngather * nchains // ngroups
        assert obs[g]["ehist"].sum() == obs[g]["nsamples"]
        assert obs[g]["D"].sum() == ngather * (nchains // ngroups) * N          # plotting.jl:115 mass check
        assert abs(obs[g]["sumE"] - sumE[g]) <= 1e-10 * max(1.0, abs(sumE[g]))
        assert abs(obs[g]["sumP"] - sumP[g]) <= 1e-10 * max(1e-6, abs(sumP[g]))


def test_full_size_invariants():
    """BASELINE config 3 (8192 chains x N=256, wall): properties that do not need an O(C N^2) CPU check.
      * running energy E0 + sum(Un-Um) equals the recomputed energy+wallsEnergy (SURVEY §4 invariant)
      * x,y stay inside the box (boundsCheck, SMC.c:529-543), histogram mass = chains*N per gather
      * identical chains with identical streams stay identical; different stream ids differ"""
    N, M, T, A = 256, 3, 1.1, 1.1
    L, Lz = geom(N)
    nchains = 8192
    orc = Oracle()
    R0, sites = orc.initialize_box(L, Lz, N)
    assert sites == N
    with smcb.Engine(nchains, N, M) as eng:
        eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=A), GOLDEN_W_M3)
        eng.broadcast_positions(R0)
        eng.set_rng(12345, 0, 0)
        eng.sweep(20, smcb.FAST)
        E, na, nt = eng.chain_state()
        ev = eng.evaluate(smcb.FAST, per_particle=False)
        R = eng.get_positions()
        eng.gather()
        obs = eng.obs_get()[0]
        tot, cut = eng.last_pair_counts()
    Erec = ev["U_lj"] + ev["U_wall"]
    assert np.all(np.abs(E - Erec) <= 1e-9 * np.maximum(1.0, np.abs(Erec)))
    assert np.all(nt == 20 * N) and np.all(na > 0) and np.all(na <= nt)
    X = R.reshape(nchains, N, 3)
    assert np.all(np.abs(X[:, :, 0]) <= L / 2) and np.all(np.abs(X[:, :, 1]) <= L / 2)
    assert obs["D"].sum() == nchains * N and obs["zprof"].sum() == nchains * N
    assert len({R[c].tobytes() for c in range(64)}) == 64          # distinct streams -> distinct chains
    acc = na.sum() / nt.sum()
    assert 0.8 < acc < 1.0                                          # ~0.94 at N=256, T=A=1.1


def test_same_stream_same_chain_and_sharding():
    """chain identity comes from (seed, global chain id): a shard starting at chain0=k reproduces
    chains k.. of the full batch exactly (this is what makes multi-GPU sharding reproducible)"""
    N, M = 108, 3
    L, Lz = geom(N)
    orc = Oracle()
    R0, _ = orc.initialize_box(L, Lz, N)
    res = {}
    for name, (cn, c0) in {"full": (8, 0), "lo": (4, 0), "hi": (4, 4)}.items():
        with smcb.Engine(cn, N, M) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=Lz), GOLDEN_W_M3)
            eng.broadcast_positions(R0)
            eng.set_rng(99, c0, 0)
            eng.sweep(10, smcb.FAST)
            res[name] = eng.get_positions()
    np.testing.assert_array_equal(res["full"][:4], res["lo"])
    np.testing.assert_array_equal(res["full"][4:], res["hi"])


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """binary checkpoint (positions, energies, counters, Philox stream position, Rbin, observable block):
    save -> destroy -> create -> load continues exactly like the uninterrupted run (the
    reference restores positions only, main.c:98-109)"""
    N, M = 108, 3
    L, Lz = geom(N)
    orc = Oracle()
    R0, _ = orc.initialize_box(L, Lz, N)
    par = smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1)

    def fresh():
        eng = smcb.Engine(6, N, M)
        eng.set_params(par, GOLDEN_W_M3)
        return eng

    with fresh() as eng:
        eng.broadcast_positions(R0)
        eng.set_rng(4242, 10, 0)
        eng.sweep(7, smcb.FAST); eng.gather()
        eng.sweep(5, smcb.FAST); eng.gather()
        ref_R, ref_state, ref_obs, ref_rbin = eng.get_positions(), eng.chain_state(), eng.obs_get()[0], eng.rbin()
    ck = tmp_path / "chains.smcb"
    with fresh() as eng:
        eng.broadcast_positions(R0)
        eng.set_rng(4242, 10, 0)
        eng.sweep(7, smcb.FAST); eng.gather()
        eng.checkpoint_save(ck)
    with fresh() as eng:
        eng.checkpoint_load(ck)
        eng.sweep(5, smcb.FAST); eng.gather()
        np.testing.assert_array_equal(eng.get_positions(), ref_R)
        for a, b in zip(eng.chain_state(), ref_state):
            np.testing.assert_array_equal(a, b)
        o = eng.obs_get()[0]
        for k in ("D", "Mu", "zprof", "ehist"):
            np.testing.assert_array_equal(o[k], ref_obs[k])
        # integer counters are exact and the moments are summed in chain order: bit-identical too
        assert o["nsamples"] == ref_obs["nsamples"]
        for k in ("sumE", "sumE2", "sumP", "sumP2", "sumAcc"):
            assert o[k] == ref_obs[k], k
        np.testing.assert_array_equal(eng.rbin(), ref_rbin)
    with smcb.Engine(5, N, M) as eng:                       # wrong shape: refused, not silently reshaped
        eng.set_params(par, GOLDEN_W_M3)
        with pytest.raises(smcb.SmcbError):
            eng.checkpoint_load(ck)


def test_obs_allreduce_single_process_two_gpus():
    """smcb_obs_allreduce: one process, one engine per GPU, NCCL all-reduce of the observable blocks - the
    path's only collective, from C.  Needs two GPUs (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, M = 108, 3
    L, Lz = geom(N)
    orc = Oracle()
    R0, _ = orc.initialize_box(L, Lz, N)
    engines, blocks = [], []
    try:
        for dev in range(2):
            eng = smcb.Engine(8, N, M, device=dev)
            eng.set_params(smcb.default_params(L=L, Lz=Lz), GOLDEN_W_M3)
            eng.broadcast_positions(R0)
            eng.set_rng(5, 8 * dev, 0)
            eng.sweep(6, smcb.FAST)
            eng.gather()
            engines.append(eng)
            blocks.append(eng.obs_get()[0])
        smcb.obs_allreduce(engines)
        for eng in engines:
            o = eng.obs_get()[0]
            for k in ("D", "Mu", "zprof", "ehist"):
                np.testing.assert_array_equal(o[k], blocks[0][k] + blocks[1][k])
            assert o["nsamples"] == 16 and o["D"].sum() == 16 * N
            assert abs(o["sumE"] - (blocks[0]["sumE"] + blocks[1]["sumE"])) <= 1e-12 * abs(o["sumE"])
    finally:
        for eng in engines:
            eng.close()


def test_every_kernel_small_case_runs_clean():
    """profiles/sanitize_case.py (written for compute-sanitizer, which is closed on this pool): every kernel and
    mode on small inputs, in a fresh process, must finish without a CUDA error"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "sanitize_case.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sanitize case done" in r.stdout, r.stderr[-1500:]


def test_full_size_run_to_run_determinism():
    """8192 chains x N=256 twice with the same streams: sweeps, all-particle steps and the gathered observable
    block are bit-identical (a data race in the warp/block-synchronised shared-memory protocols would show here)"""
    N, M = 256, 3
    L, Lz = geom(N)
    orc = Oracle()
    R0, _ = orc.initialize_box(L, Lz, N)
    runs = []
    for _ in range(2):
        with smcb.Engine(8192, N, M) as eng:
            eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=1.1), GOLDEN_W_M3)
            eng.broadcast_positions(R0)
            eng.set_rng(777, 0, 0)
            eng.sweep(6, smcb.FAST)
            eng.gather()
            Rs, Es = eng.get_positions().copy(), eng.chain_state()[0].copy()
            eng.set_params(smcb.default_params(L=L, Lz=Lz, T=1.1, A=2e-4), GOLDEN_W_M3)
            eng.step_allparticle(6, smcb.FAST)
            eng.gather()
            o = eng.obs_get()[0]
            runs.append((Rs, Es, eng.get_positions().copy(), eng.chain_state()[0].copy(), o["D"].copy(), o["sumE"], o["sumP"]))
    for a, b in zip(runs[0], runs[1]):
        np.testing.assert_array_equal(a, b)
