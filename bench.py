#!/usr/bin/env python
"""bench.py — throughput of the Smart-MC hot path on B200, with the FP64 roofline and the CPU baseline.

Workload (BASELINE.json configs[2]): 8192 independent chains x N=256 LJ molecules above the wall,
main.c geometry (L=33, Lz=240, T=1.1, A=T, M=3), fcc start lattice of initializeBox, PER GPU
(weak scaling: chains shard across ranks with no data-path collective; only the observable block
is all-reduced).  One "step" = `sweeps_per_step` sweeps (oneParticleMoves, SMC.c:278-351) of every
chain in one kernel launch, followed by one gather of the observables (+ the NCCL all-reduce of the
observable block when N>1).  sweeps_per_step defaults to 40 = the reference's own gather cadence
(main.c:15-18: 16e6 steps / 4e5 samples).

Metric: pair-interactions/s (one ORDERED (i,j) geometry + LJ energy + force evaluation; a sweep is
2*N*(N-1) of them); chain-steps/s (sweeps/s) is reported beside it.

  value     device-resident throughput, timed with CUDA events on the engine's launch stream
  e2e       the same through the public C-ABI with HOST buffers: upload positions, sweeps, gather,
            download positions + chain state, every step
  roofline  FP64-pipe bound (the pair kernel is compute bound: ~3.3 MFLOP per 12 KB of state):
            achieved = (17*pairs + 16*pairs_in_cutoff) flops / kernel time, peak = DFMA peak
            measured live by smcb_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, compiled from
            /root/reference by oracle/build_ref.sh) running oneParticleMoves on all host cores,
            one independent chain per core - the only parallel mode the reference has.
"""
import argparse
import ctypes
import importlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PART, M_SITES, L_BOX, LZ_BOX, TEMP = 256, 3, 33.0, 240.0, 1.1
FLOPS_PAIR, FLOPS_INCUT = 17.0, 16.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smcb200", choices=["smcb200", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default 8192; largeN: 256/world)")
    ap.add_argument("--sweeps-per-step", type=int, default=40)
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--kernel", default="sweep", choices=["sweep", "allparticle"])
    ap.add_argument("--workload", default="batched", choices=["batched", "bulk", "grid", "largeN"],
                    help="batched: 8192 chains x N=256 per GPU (configs[2], the headline); grid: 65536 chains in total on a "
                         "16 T x 4 Lz x 4 wall-strength grid x 256 replicas, sharded over the GPUs, one observable group per "
                         "grid point (configs[3]); largeN: 256 chains x N=4096 in total, sharded over the GPUs (configs[4], "
                         "all-particle kernel with thread-block clusters); bulk: 8192 chains x N=108 per GPU, 3-D periodic, "
                         "rho*=0.5, T*=1.0, cutoff L/2 (configs[0], the SMC_noMPI_noWall geometry)")
    ap.add_argument("--largeN-sweep", action="store_true",
                    help="largeN workload with the reference's sweep (block-per-chain kernel, N > 512) instead of the all-particle step")
    ap.add_argument("--start", default="lattice", choices=["lattice", "droplet"],
                    help="lattice: initializeBox's dilute fcc lattice (the reference's start); droplet: all molecules condensed "
                         "on the lower wall (jittered simple-cubic block, spacing 1.12: ~80 partners inside the cutoff each) - "
                         "the state long reference runs end in")
    ap.add_argument("--thermalise", type=int, default=2000,
                    help="sweeps with 2A (sMC's thermalisation, SMC.c:110-125) before the extra 'thermalised' timing leg; 0 = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU reference arm
def _ref_worker(args):
    """one host core = one independent chain of the compiled reference (oracle/_ref)"""
    libname, nsweeps, seed = args
    from oracle_bindings import GOLDEN_W_M3, RefLib
    ref = RefLib(N_PART, M_SITES, fast=libname.endswith("_fast"))
    R = ref.initializeBox(L_BOX, LZ_BOX)
    W = GOLDEN_W_M3.copy()
    ref.lib.oracle_srand(seed)
    Rn = np.zeros_like(R)
    j = ctypes.c_int(0)
    e = ctypes.c_double(0.0)
    t0 = time.perf_counter()
    for _ in range(nsweeps):
        ref.lib.oneParticleMoves(R, Rn, W, L_BOX, LZ_BOX, TEMP, TEMP, ctypes.byref(j), ctypes.byref(e))
    return time.perf_counter() - t0, j.value


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_steps(pool, cores, nsweeps, nsteps, fast=False):
    """nsteps x (every core runs nsweeps sweeps of its own chain); returns wall seconds per step"""
    lib = "ref_fast" if fast else "ref"
    times = []
    for s in range(nsteps):
        t0 = time.perf_counter()
        pool.map(_ref_worker, [(lib, nsweeps, 1000 * s + c) for c in range(cores)])
        times.append(time.perf_counter() - t0)
    return times


def pairs_per_sweep(n):
    return 2.0 * n * (n - 1)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_so = os.path.join(ROOT, "oracle", "_ref", f"libref_N{N_PART}_M{M_SITES}.so")
    if not os.path.exists(ref_so):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return
    cores = host_cores()
    nsweeps = max(1, min(args.sweeps_per_step, 40))
    with mp.get_context("fork").Pool(cores) as pool:
        run_reference_steps(pool, cores, 2, max(1, min(args.warmup, 3)))
        times = run_reference_steps(pool, cores, nsweeps, args.steps)
    total = sum(times)
    sweeps = cores * nsweeps * args.steps
    value = sweeps * pairs_per_sweep(N_PART) / total
    line = {
        "impl": "reference", "metric": "pair_interactions_per_s", "value": value, "unit": "pair-interactions/s",
        "chain_steps_per_s": sweeps / total, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"reference oneParticleMoves (SMC.c:278-351), N={N_PART} wall, one chain per host core",
                   "N": N_PART, "M": M_SITES, "L": L_BOX, "Lz": LZ_BOX, "T": TEMP, "A": TEMP,
                   "sweeps_per_step": nsweeps, "chains": cores},
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": cores, "kind": "reference",
                         "sample": f"{cores} chains x {nsweeps} sweeps x {args.steps} steps, gcc -O2 -ffp-contract=off build of /root/reference"},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args):
    """bounded sample of the reference on all host cores (rank 0, N=1 only)"""
    ref_so = os.path.join(ROOT, "oracle", "_ref", f"libref_N{N_PART}_M{M_SITES}.so")
    if not os.path.exists(ref_so):
        return None
    cores = host_cores()
    out = {}
    with mp.get_context("fork").Pool(cores) as pool:
        for fast in (False, True):
            run_reference_steps(pool, cores, 2, 1, fast)
            t_probe = run_reference_steps(pool, cores, 10, 1, fast)[0]
            nsweeps = int(max(10, min(4000, 10 * (args.cpu_seconds / 2) / max(t_probe, 1e-3))))
            t = run_reference_steps(pool, cores, nsweeps, 1, fast)[0]
            out["fast" if fast else "parity"] = (cores * nsweeps / t, nsweeps, t)
    sps, nsw, t = out["parity"]
    return {"value": sps * pairs_per_sweep(N_PART), "unit": "pair-interactions/s", "cores": cores, "kind": "reference",
            "chain_steps_per_s": sps, "chain_steps_per_s_per_core": sps / cores,
            "sample": f"{cores} host cores x 1 chain x {nsw} sweeps of oneParticleMoves (N={N_PART}, wall), {t:.1f} s, "
                      "reference compiled -O2 -ffp-contract=off",
            "courtesy_O3_avx2_value": out["fast"][0] * pairs_per_sweep(N_PART)}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    smcb = importlib.import_module("montecarlo-surfacer_b200")
    GOLDEN_W_M3 = smcb.REFERENCE_WALL_M3          # main.c's wall table; nothing under oracle/ is touched by this arm

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; smcb200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    Cn, N, S = args.chains or 8192, N_PART, args.sweeps_per_step
    total_largeN = 256
    if args.workload == "largeN":                   # configs[4]: 256 chains x N=4096 in total, strong-sharded
        N, args.kernel = 4096, ("sweep" if args.largeN_sweep else "allparticle")
        Cn = args.chains or smcb.shard_chains(256, world, rank).nchains      # --chains 32 emulates one rank of 8
        total_largeN = float(world) * Cn if args.chains else 256.0
    bulk = args.workload == "bulk"
    if bulk:                                        # configs[0]: SMC_noMPI_noWall.c:74-143 geometry, fcc start (its :359-394)
        N = 108
    grid = None
    if args.workload == "grid":                     # configs[3]: the T / density / wall grid the MPI ranks used to split
        total_grid = (args.chains * world) if args.chains else 65536
        shard = smcb.shard_chains(total_grid, world, rank)
        Cn = shard.nchains
        temps = [0.7 + 0.05 * i for i in range(16)]
        lzs = [120.0, 160.0, 200.0, 240.0]
        walls = [0, 1, 2, 3]                         # wall tables: ymin = 2.0, 2.67, 3.33, 4.0 (main.c:76 uses 3.0 +- 0.5)
        grid = (shard, temps, lzs, walls, total_grid)
    mode = smcb.STRICT if args.mode == "strict" else smcb.FAST
    A = TEMP if args.kernel == "sweep" else (2e-4 if N <= 256 else 2e-6)
    if args.start == "droplet" and args.kernel == "sweep":
        A = 0.02                                     # A = T moves 1.5 sigma per trial: nothing is accepted inside a liquid

    # start lattice of initializeBox(33, 240, 256) (SMC.c:413-465): 4x4x4 fcc cells, a = 8.25, shifted a/4
    # (N=4096: 16x16x4 cells, a = 33/16 - the reference's own generator is invalid there)
    nxy, nz = (4, 4) if N == 256 else ((3, 3) if N == 108 else (16, 4))
    Lb = (108 / 0.5) ** (1.0 / 3.0)                 # bulk box: rho* = 0.5
    a = (Lb if bulk else L_BOX) / nxy
    cells = np.array([(i, j, k) for i in range(nxy) for j in range(nxy) for k in range(nz)], dtype=float)
    basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    X = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + a / 4
    Pz = LZ_BOX - LZ_BOX / 20.0
    if bulk:
        X -= Lb * np.rint(X / Lb)
    else:
        X[:, :2] -= L_BOX * np.rint(X[:, :2] / L_BOX)
        X[:, 2] -= Pz * np.rint(X[:, 2] / Pz)
    R0 = X.reshape(-1)
    if args.start == "droplet":
        nzl = 4 if N <= 512 else 8
        nxy = int(np.ceil(np.sqrt(N / nzl)))
        g = np.array([(i, j, k) for k in range(nzl) for i in range(nxy) for j in range(nxy)], dtype=float)[:N]
        g[:, 0] = (g[:, 0] - nxy / 2) * 1.12
        g[:, 1] = (g[:, 1] - nxy / 2) * 1.12
        g[:, 2] = -LZ_BOX / 2 + 0.95 + g[:, 2] * 1.12
        rs = np.random.default_rng(7)
        g += (rs.random(g.shape) * 2 - 1) * 0.05
        R0 = g[rs.permutation(N)].reshape(-1)

    eng = smcb.Engine(Cn, N, M_SITES, device=local)
    if grid:
        shard, temps, lzs, walls, total_grid = grid
        params, ngroups = smcb.grid_chain_params(shard, temps, lzs, walls, L=L_BOX)
        Wt = np.concatenate([GOLDEN_W_M3 * (ym / 3.0) for ym in (2.0, 8.0 / 3.0, 10.0 / 3.0, 4.0)])
        eng.set_params(params, Wt, ngroups=ngroups)
        chain0 = shard.chain0
    elif bulk:
        A = 0.01 if args.kernel == "sweep" else 1e-4      # a liquid: the prototype's own A is 4e-8 (SMC_noMPI_noWall.c:192)
        eng.set_params(smcb.default_params(L=Lb, Lz=Lb, T=1.0, A=A, rc2=Lb * Lb / 4, flags=smcb.PERIODIC_Z), None, ngroups=1)
        chain0 = rank * Cn
    else:
        eng.set_params(smcb.default_params(L=L_BOX, Lz=LZ_BOX, T=TEMP, A=A), GOLDEN_W_M3, ngroups=1)
        chain0 = rank * Cn
    eng.obs_configure(nebins=64, e_lo=-8.0, e_hi=2.0)
    eng.broadcast_positions(R0)
    eng.set_rng(12345, chain0, 0)
    info = eng.device_info()
    lay = eng.obs_layout()

    obs_cnt = torch.zeros(lay.u64_total, dtype=torch.int64, device="cuda")
    obs_mom = torch.zeros(lay.f64_total, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def run_kernel(kernel):
        if kernel == "sweep":
            eng.sweep(S, mode)
        else:
            eng.step_allparticle(S, mode)
        return eng.last_kernel_ms()[0]

    def allreduce_obs():
        """the ONLY collective of the path: sum the observable block over ranks (NCCL)"""
        if world == 1:
            return 0.0
        eng.obs_export_device(obs_cnt.data_ptr(), obs_mom.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smcb.allreduce_observables(obs_cnt, obs_mom)
        e1.record()
        torch.cuda.synchronize()
        eng.obs_reset()          # ranks keep accumulating deltas; the reduced block lives in obs_cnt/obs_mom
        return e0.elapsed_time(e1)

    def one_step(kernel):
        k_ms = run_kernel(kernel)
        pairs = eng.last_pair_counts()
        eng.gather()
        g_ms = eng.last_kernel_ms()[0]
        c_ms = allreduce_obs()
        return k_ms, g_ms, c_ms, pairs

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_leg(kernel, nsteps, warmup, sample_clocks=False):
        """`warmup` untimed steps, then exactly `nsteps` steps between barriers; device times (CUDA events
        on the engine's stream) summed per rank, max over ranks"""
        for _ in range(warmup):
            one_step(kernel)
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        k_tot = g_tot = c_tot = 0.0
        pairs_tot = pairs_cut = 0
        eng.reset_counters()
        for _ in range(nsteps):
            flush.fill_(1)                      # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize()
            k_ms, g_ms, c_ms, pairs = one_step(kernel)
            k_tot += k_ms; g_tot += g_ms; c_tot += c_ms
            pairs_tot += pairs[0]; pairs_cut += pairs[1]
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        _, na, nt = eng.chain_state()
        dev_ms = smcb.max_over_ranks(k_tot + g_tot + c_tot, device="cuda")
        k_max = smcb.max_over_ranks(k_tot, device="cuda")
        unit_pairs = pairs_per_sweep(N) if kernel == "sweep" else float(N) * (N - 1)
        total_chains = float(world) * Cn if args.workload in ("batched", "bulk") else (float(grid[4]) if grid else total_largeN)
        chain_steps = total_chains * S * nsteps
        flops = FLOPS_PAIR * pairs_tot + FLOPS_INCUT * pairs_cut
        return {"value": chain_steps * unit_pairs / (dev_ms * 1e-3), "chain_steps_per_s": chain_steps / (dev_ms * 1e-3),
                "ms_per_step": dev_ms / nsteps, "kernel_ms_per_step": k_max / nsteps, "gather_ms_per_step": g_tot / nsteps,
                "allreduce_ms_per_step": c_tot / nsteps, "wall_s": wall, "clocks": clocks,
                "pairs_in_cutoff_frac": pairs_cut / max(1, pairs_tot), "acceptance": float(na.sum()) / max(1, int(nt.sum())),
                "achieved_tflops": flops / (k_tot * 1e-3) / 1e12, "unit_pairs": unit_pairs}

    fp64_peak, _ = eng.measure_fp64_peak()
    W = max(args.warmup, 3)
    main = timed_leg(args.kernel, args.steps, W, sample_clocks=True)

    # ---- end to end through the C-ABI with host buffers (every rank, max over ranks) ----------
    e2e = None
    if not args.no_e2e:
        host_R = torch.empty((Cn, 3 * N), dtype=torch.float64).pin_memory().numpy()
        nsteps_e2e = max(3, min(args.steps, 20))            # the same window of the trajectory as the device-resident leg

        def e2e_step():
            eng.set_positions(host_R)                  # H2D of the step's inputs
            run_kernel(args.kernel)
            eng.gather()
            allreduce_obs()
            eng.get_positions(host_R)                  # D2H of the step's results
            return eng.chain_state()

        # same start as the device-resident leg: the lattice, the same streams, W warm-up steps
        eng.broadcast_positions(R0)
        eng.set_rng(12345, chain0, 0)
        eng.get_positions(host_R)
        for _ in range(W):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps_e2e):
            e2e_step()
        barrier()
        te = smcb.max_over_ranks(time.perf_counter() - t0, device="cuda")
        total_chains = float(world) * Cn if args.workload in ("batched", "bulk") else (float(grid[4]) if grid else total_largeN)
        e2e = {"value": total_chains * S * nsteps_e2e * main["unit_pairs"] / te, "unit": "pair-interactions/s",
               "chain_steps_per_s": total_chains * S * nsteps_e2e / te, "steps": nsteps_e2e,
               "h2d_bytes_per_step": int(Cn * 3 * N * 8), "d2h_bytes_per_step": int(Cn * 3 * N * 8 + Cn * 24),
               "timing": "host wall clock around set_positions -> kernel -> gather -> get_positions -> chain_state; same start "
                         "(lattice + warm-up steps) as the device-resident leg, no L2 flush between steps"}

    # ---- extra legs (reported beside the headline, same JSON line) -----------------------------
    extra = {}
    if args.workload == "batched" and args.kernel == "sweep" and mode == smcb.FAST and args.start == "lattice":
        if args.thermalise > 0:
            # sMC's thermalisation (2A, SMC.c:110-125) so molecules reach the wall and partners become common
            eng.set_step_scale(2.0)
            eng.sweep(args.thermalise, mode)
            eng.set_step_scale(1.0)
            th = timed_leg("sweep", max(2, min(args.steps, 3)), 1)
            extra["thermalised"] = {k: th[k] for k in ("value", "chain_steps_per_s", "ms_per_step", "kernel_ms_per_step",
                                                       "pairs_in_cutoff_frac", "acceptance")}
            extra["thermalised"].update(sweeps_before=args.thermalise, roofline_frac=th["achieved_tflops"] / fp64_peak)
        # north-star kernel B on the same chains: all-particle steps need a small A to be accepted at all
        eng.set_params(smcb.default_params(L=L_BOX, Lz=LZ_BOX, T=TEMP, A=2e-4), GOLDEN_W_M3, ngroups=1)
        eng.obs_configure(nebins=64, e_lo=-8.0, e_hi=2.0)
        eng.broadcast_positions(R0)
        ap_ = timed_leg("allparticle", max(2, min(args.steps, 3)), 2)
        extra["allparticle_kernel"] = {k: ap_[k] for k in ("value", "chain_steps_per_s", "ms_per_step", "kernel_ms_per_step",
                                                            "pairs_in_cutoff_frac", "acceptance")}
        extra["allparticle_kernel"].update(A=2e-4, roofline_frac=ap_["achieved_tflops"] / fp64_peak,
                                           note="one all-particle Smart-MC step = N(N-1) ordered pair-interactions")

    if rank == 0:
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(args.kernel)
            if tj:
                traffic = tj["dram_bytes_per_chain_launch"] * Cn
                traffic_note = tj["source"]
        line = {
            "metric": "pair_interactions_per_s", "value": main["value"], "unit": "pair-interactions/s",
            "chain_steps_per_s": main["chain_steps_per_s"],
            "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if args.workload in ("batched", "bulk") or (grid and args.chains) else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"{Cn} chains/GPU x N={N} with wall (BASELINE configs[2])" if args.workload == "batched"
                                    else f"{Cn} chains/GPU x N={N} bulk, 3-D periodic, rho*=0.5, T*=1.0 (BASELINE configs[0] geometry)" if bulk
                                    else f"{grid[4]} chains x N={N} on a 16 T x 4 Lz x 4 wall grid x replicas (BASELINE configs[3]), "
                                         f"{Cn} on this rank, {len(grid[1]) * len(grid[2]) * len(grid[3])} observable groups all-reduced"
                                    if grid else f"{int(total_largeN)} chains x N={N} with wall in total (BASELINE configs[4]), {Cn} on this rank")
                                   + f", {args.kernel} kernel, {args.mode}",
                       "chains_per_gpu": Cn, "N": N, "M": M_SITES, "L": Lb if bulk else L_BOX, "Lz": Lb if bulk else LZ_BOX,
                       "T": 1.0 if bulk else TEMP, "A": A,
                       "sweeps_per_step": S, "start": ("initializeBox fcc lattice" if args.start == "lattice" else "condensed droplet on the wall") + " + warm-up steps",
                       "l2": "flushed between timed steps (256 MB write)", "rng": "Philox4x32-10",
                       "parallelism": f"chains sharded x{world}, NCCL all-reduce of the observable block only"},
            "kernel_ms_per_step": main["kernel_ms_per_step"], "gather_ms_per_step": main["gather_ms_per_step"],
            "allreduce_ms_per_step": main["allreduce_ms_per_step"], "wall_s_timed_region": main["wall_s"],
            "pairs_in_cutoff_frac": main["pairs_in_cutoff_frac"], "acceptance": main["acceptance"],
            "roofline": {"bound": "fp64", "achieved": main["achieved_tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": main["achieved_tflops"] / fp64_peak, "traffic": traffic, "traffic_note": traffic_note,
                         "note": "ALGORITHMIC flops = 17 per ordered pair-interaction + 16 more inside the cutoff (SURVEY §8d), "
                                 "a sweep counted as the reference executes it: 2N(N-1) pair-interactions (old + proposed "
                                 "position of every trial).  The kernel evaluates fewer: old-position terms are cached, and the "
                                 "cutoff screen runs in packed FP32 (exact FP64 for the pairs inside), so frac is a figure of "
                                 "merit against the FP64 peak, not FP64-pipe utilisation - that is in profiles/ (ncu).  peak = "
                                 "DFMA peak measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)"},
            "clocks": main["clocks"], "gpu_launches": args.steps * 4, "device": info,
        }
        line.update(extra)
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline and N == N_PART:
            cb = cpu_baseline(args)
            if cb:
                line["cpu_baseline"] = cb
        print(json.dumps(line), file=real_stdout, flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
