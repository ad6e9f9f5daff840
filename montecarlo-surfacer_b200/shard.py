"""Chain sharding over the GPUs of one box (one process per GPU, torch.distributed).

The reference runs one chain per process and its ranks never talk (`rank` only
names the CSV files, SMC.c:66-95).  Here the chain grid - the temperature / density (Lz) /
wall-strength points the MPI ranks used to split, times replicas - is cut into contiguous blocks,
one block per rank, with the grid points INTERLEAVED in chain order so that every rank holds every
grid point (condensed chains accept differently but cost the same per sweep, so the blocks stay
balanced).  There is no data-path collective: the ONLY exchange is the sum of the packed observable
block (voxel density + mobility counts, z profile, energy histogram, moments; smcb_obs_layout) over
ranks at a gather, one all-reduce of integers and one of doubles.
"""
from dataclasses import dataclass

from .engine import default_params


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    chain0: int      # global id of the first chain of this rank (the Philox stream id base)
    nchains: int     # chains on this rank
    total: int       # chains in the whole job


def shard_chains(total, world, rank):
    """contiguous blocks; the first `total % world` ranks take one extra chain"""
    if not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    n = base + (1 if rank < extra else 0)
    c0 = rank * base + min(rank, extra)
    return Shard(rank, world, c0, n, total)


def grid_points(temps, lzs, wall_ids):
    """the parameter grid, group id = index into the returned list"""
    return [(float(T), float(Lz), int(w)) for T in temps for Lz in lzs for w in wall_ids]


def grid_chain_params(shard, temps, lzs, wall_ids, L=33.0, gamma=1.0, **kw):
    """ChainParams for the chains [chain0, chain0+nchains) of a job whose global chain g sits on grid
    point g % npoints (interleaved) and is replica g // npoints; A = gamma*T (main.c:48-51)."""
    pts = grid_points(temps, lzs, wall_ids)
    out = []
    for g in range(shard.chain0, shard.chain0 + shard.nchains):
        gi = g % len(pts)
        T, Lz, w = pts[gi]
        out.append(default_params(L=L, Lz=Lz, T=T, A=gamma * T, wall=w, group=gi, **kw))
    return out, len(pts)


def allreduce_observables(counters, moments, group=None):
    """sum the packed observable block over ranks, in place.  `counters` is an int64 view of the
    uint64 counter block (torch has no uint64 reduction; counts stay far below 2^63) and `moments`
    a float64 tensor; both live on the device for NCCL, on the host for gloo."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)


def max_over_ranks(value, device=None, group=None):
    """max of a python float over ranks (timing rule: device times are combined as the max)"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


__all__ = ["Shard", "shard_chains", "grid_points", "grid_chain_params", "allreduce_observables", "max_over_ranks"]
