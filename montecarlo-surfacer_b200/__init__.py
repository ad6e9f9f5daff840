"""smcb200 — host-side binding of libsmcb200.so, the B200-native Smart-Monte-Carlo engine.

The product is the CUDA shared library (csrc/, C ABI in include/smcb200.h) and the C drop-in
sources in dropin/ that re-host the reference's SMC.h API on it.  This package is only the ctypes
view of that ABI used by tests/, bench.py and Python drivers.  It never falls back to a CPU
implementation: importing works without a GPU (so the C-ABI export check can run), creating an
Engine does not.
"""
from .engine import (FAST, STRICT, FP32, WALL, PERIODIC_Z, ChainParams, Engine, ObsLayout, SmcbError,
                     default_params, lib_path, load_library, exported_symbols, header_symbols,
                     obs_layout_host, unpack_obs, obs_allreduce, REFERENCE_WALL_M3)
from .shard import (Shard, shard_chains, grid_points, grid_chain_params, allreduce_observables,
                    max_over_ranks)

__all__ = ["FAST", "STRICT", "FP32", "WALL", "PERIODIC_Z", "ChainParams", "Engine", "ObsLayout", "SmcbError",
           "default_params", "lib_path", "load_library", "exported_symbols", "header_symbols",
           "obs_layout_host", "unpack_obs", "obs_allreduce", "REFERENCE_WALL_M3", "Shard", "shard_chains", "grid_points", "grid_chain_params",
           "allreduce_observables", "max_over_ranks"]
