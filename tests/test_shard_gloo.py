"""Host-side multi-GPU logic on CPU: chain sharding and the ONE collective of the path, the sum of
the packed observable block over ranks, run with torch.distributed `gloo`, world_size 2.
Each rank builds the block of its shard with the oracle's localDensityAndMobility restatement
(SMC.c:912-927) - the same layout libsmcb200's k_gather fills - all-reduces it, and the result
must equal the block of the unsharded job."""
import importlib
import os
import socket

import numpy as np
import pytest

from oracle_bindings import Oracle, config_gas, make_sys

smcb = importlib.import_module("montecarlo-surfacer_b200")


def test_shard_partition_is_exact():
    for total in (0, 1, 7, 8192, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            shards = [smcb.shard_chains(total, world, r) for r in range(world)]
            assert shards[0].chain0 == 0
            for a, b in zip(shards, shards[1:]):
                assert b.chain0 == a.chain0 + a.nchains
            assert shards[-1].chain0 + shards[-1].nchains == total
            assert max(s.nchains for s in shards) - min(s.nchains for s in shards) <= 1
    with pytest.raises(ValueError):
        smcb.shard_chains(10, 2, 2)


def test_grid_is_interleaved_over_ranks():
    temps, lzs, walls = [0.7, 0.9, 1.1, 1.3], [120.0, 240.0], [0, 1]
    npts = len(temps) * len(lzs) * len(walls)
    total = npts * 8
    seen = []
    for r in range(4):
        sh = smcb.shard_chains(total, 4, r)
        params, ng = smcb.grid_chain_params(sh, temps, lzs, walls)
        assert ng == npts and len(params) == sh.nchains
        groups = [p.group for p in params]
        assert set(groups) == set(range(npts))            # every rank holds every grid point
        for p in params:
            T, Lz, w = smcb.grid_points(temps, lzs, walls)[p.group]
            assert (p.T, p.Lz, p.wall, p.A) == (T, Lz, w, T)   # A = gamma*T, main.c:48-51
        seen += groups
    assert np.bincount(seen).tolist() == [8] * npts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _block_of(orc, lay, R, groups, N, L, Lz, energies):
    """host restatement of what smcb_gather adds for chains R[c] (group ids `groups`)"""
    cnt = np.zeros(lay.u64_total, dtype=np.uint64)
    mom = np.zeros(lay.f64_total)
    s = make_sys(N, 3, L, Lz)
    nv, nz, ne = lay.nvox, lay.nz, lay.nebins
    for c in range(R.shape[0]):
        base = groups[c] * lay.u64_per_group
        D = np.zeros(nv, dtype=np.uint64); Mu = np.zeros(nv, dtype=np.uint64); Rbin = np.zeros(N, dtype=np.int32)
        orc.local_density(s, np.ascontiguousarray(R[c]), D, Rbin, Mu)
        cnt[base:base + nv] += D
        cnt[base + nv:base + 2 * nv] += Mu
        cnt[base + 2 * nv:base + 2 * nv + nz] += D.reshape(33, 33, 33).sum(axis=(0, 1))
        b = int(np.clip(np.floor((energies[c] / N - lay.e_lo) / (lay.e_hi - lay.e_lo) * ne), 0, ne - 1))
        cnt[base + 2 * nv + nz + b] += 1
        cnt[base + 2 * nv + nz + ne] += 1
        m = groups[c] * lay.f64_per_group
        mom[m] += energies[c]
        mom[m + 1] += energies[c] ** 2
    return cnt, mom


def _job(total=12, N=32, L=33.0, Lz=200.0, ngroups=3):
    rng = np.random.default_rng(5)
    R = np.stack([config_gas(N, L, Lz, rng) for _ in range(total)])
    groups = [g % ngroups for g in range(total)]
    energies = rng.standard_normal(total) * 10 - 50
    return R, groups, energies


def _rank_main(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        R, groups, energies = _job()
        lay = smcb.obs_layout_host(3, nebins=16, e_lo=-8.0, e_hi=2.0)
        sh = smcb.shard_chains(R.shape[0], world, rank)
        sl = slice(sh.chain0, sh.chain0 + sh.nchains)
        cnt, mom = _block_of(orc, lay, R[sl], groups[sl], 32, 33.0, 200.0, energies[sl])
        tc = torch.from_numpy(cnt.view(np.int64).copy())
        tm = torch.from_numpy(mom.copy())
        smcb.allreduce_observables(tc, tm)
        tmax = smcb.max_over_ranks(1.0 + rank)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cnt=tc.numpy().view(np.uint64), mom=tm.numpy(), tmax=tmax)
    finally:
        dist.destroy_process_group()


def test_observable_allreduce_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    orc = Oracle()
    R, groups, energies = _job()
    lay = smcb.obs_layout_host(3, nebins=16, e_lo=-8.0, e_hi=2.0)
    cnt, mom = _block_of(orc, lay, R, groups, 32, 33.0, 200.0, energies)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        np.testing.assert_array_equal(got["cnt"], cnt)            # integer counters: exact
        np.testing.assert_allclose(got["mom"], mom, rtol=1e-14)    # doubles: summation order only
        assert float(got["tmax"]) == 2.0
    grp = smcb.unpack_obs(lay, cnt, mom)
    assert sum(int(g["D"].sum()) for g in grp) == R.shape[0] * 32
    assert [g["nsamples"] for g in grp] == [4, 4, 4]
