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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="smcb200", choices=["smcb200", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default 8192; largeN: 256/world)")
    ap.add_argument("--sweeps-per-step", type=int, default=40)
    ap.add_argument("--mode", default="fast", choices=["fast", "strict", "fp32"],
                    help="fast / strict: FP64 arithmetic (fast = the headline); fp32: the optional single-precision mode "
                         "(all-particle kernel only; its own line, \"dtype\": \"f32\", never the headline)")
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
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (thermalised gas, all-particle kernel, droplet)")
    ap.add_argument("--tune-step", type=float, default=0.0,
                    help="before the timed legs, adapt every chain's A to this acceptance (smcb_tune_step_size); 0 = keep the configured A")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="batched/bulk workloads: weak = --chains (8192) chains PER GPU; strong = that many chains IN TOTAL, "
                         "sharded over the GPUs (north_star: 1->8 GPUs on 8192 chains x N=256)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU reference arm
# One PERSISTENT worker process per host core, each owning one independent chain of the compiled reference
# (oracle/_ref) - the only parallel mode the reference has.  The library is loaded and the start lattice built
# once, outside every timed window; a step is "every worker runs nsweeps sweeps", timed INSIDE the worker around
# the oneParticleMoves loop (SMC.c:278-351); the step's time is the slowest worker's.
def _ref_worker_main(conn, core, fast, seed):
    try:
        try:
            os.sched_setaffinity(0, {core})
        except (AttributeError, OSError):
            pass
        from oracle_bindings import GOLDEN_W_M3, RefLib
        ref = RefLib(N_PART, M_SITES, fast=fast)
        R = ref.initializeBox(L_BOX, LZ_BOX)
        W = GOLDEN_W_M3.copy()
        ref.lib.oracle_srand(seed)
        Rn = np.zeros_like(R)
        j = ctypes.c_int(0)
        e = ctypes.c_double(0.0)
        conn.send("ready")
        while True:
            nsweeps = conn.recv()
            if nsweeps is None:
                break
            t0 = time.perf_counter()
            for _ in range(nsweeps):
                ref.lib.oneParticleMoves(R, Rn, W, L_BOX, LZ_BOX, TEMP, TEMP, ctypes.byref(j), ctypes.byref(e))
            conn.send((time.perf_counter() - t0, j.value))
    except Exception as ex:                       # surfaces in the parent instead of a silent hang
        conn.send(("error", repr(ex)))


def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


class RefWorkers:
    def __init__(self, fast=False):
        ctx = mp.get_context("fork")
        self.cores = host_cores()
        self.procs, self.conns = [], []
        for i, core in enumerate(self.cores):
            pc, cc = ctx.Pipe()
            pr = ctx.Process(target=_ref_worker_main, args=(cc, core, fast, 1000 + i), daemon=True)
            pr.start()
            self.procs.append(pr); self.conns.append(pc)
        for c in self.conns:
            msg = c.recv()
            if msg != "ready":
                raise RuntimeError(f"reference worker failed: {msg}")

    def step(self, nsweeps):
        """every worker runs nsweeps sweeps; returns the slowest worker's in-loop seconds"""
        for c in self.conns:
            c.send(nsweeps)
        out = [c.recv() for c in self.conns]
        for o in out:
            if o[0] == "error":
                raise RuntimeError(f"reference worker failed: {o[1]}")
        return max(o[0] for o in out)

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except (BrokenPipeError, OSError):
                pass
        for pr in self.procs:
            pr.join(timeout=5)


def pairs_per_sweep(n):
    return 2.0 * n * (n - 1)


def run_reference(nsteps, warmup, min_step_s=0.25, fast=False):
    """(value pair-interactions/s, cores, sweeps per step, total timed seconds): `warmup` untimed steps, then nsteps
    timed ones of at least min_step_s each (a step of 40 sweeps is 80 ms of CPU work: too short to time from outside)"""
    w = RefWorkers(fast)
    try:
        w.step(40)                                          # cold start (page faults, frequency ramp): not a calibration
        # warm up until the host has settled: 40-sweep steps until two in a row agree within 3 % (at most ~3 s) - on a
        # fresh box the first second runs 15-50 % slower, and a slow reference would flatter the GPU arm
        t40, spent = w.step(40), 0.0
        for _ in range(40):
            t = w.step(40)
            spent += t
            settled = abs(t - t40) <= 0.03 * t
            t40 = t
            if settled or spent > 3.0:
                break
        nsweeps = int(max(40, np.ceil(40 * min_step_s / max(t40, 1e-4))))
        for _ in range(max(1, warmup - 1)):
            w.step(nsweeps)
        total = sum(w.step(nsweeps) for _ in range(nsteps))
    finally:
        w.close()
    cores = len(w.cores)
    return cores * nsweeps * nsteps * pairs_per_sweep(N_PART) / total, cores, nsweeps, total


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_so = os.path.join(ROOT, "oracle", "_ref", f"libref_N{N_PART}_M{M_SITES}.so")
    if not os.path.exists(ref_so):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return
    nsteps = max(1, min(args.steps, 40))
    value, cores, nsweeps, total = run_reference(nsteps, max(1, min(args.warmup, 3)), min_step_s=0.5)
    line = {
        "impl": "reference", "metric": "pair_interactions_per_s", "value": value, "unit": "pair-interactions/s",
        "chain_steps_per_s": value / pairs_per_sweep(N_PART), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"reference oneParticleMoves (SMC.c:278-351), N={N_PART} with wall (BASELINE configs[2] chains), "
                               "one chain per host core", "N": N_PART, "M": M_SITES, "L": L_BOX, "Lz": LZ_BOX, "T": TEMP, "A": TEMP,
                   "sweeps_per_step": nsweeps, "chains": cores, "start": "initializeBox fcc lattice",
                   "timing": "inside each persistent worker around its sweep loop; a step lasts as long as its slowest worker"},
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": cores, "kind": "reference",
                         "sample": f"{cores} chains x {nsweeps} sweeps x {nsteps} steps ({total:.1f} s timed), "
                                   "gcc -O2 -ffp-contract=off build of /root/reference"},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args):
    """bounded sample of the reference on all host cores (rank 0, N=1 only): the same measurement as --impl reference"""
    ref_so = os.path.join(ROOT, "oracle", "_ref", f"libref_N{N_PART}_M{M_SITES}.so")
    if not os.path.exists(ref_so):
        return None
    nsteps = max(4, int(args.cpu_seconds / 2 / 0.5))
    value, cores, nsweeps, total = run_reference(nsteps, 2, min_step_s=0.5)
    out = {"value": value, "unit": "pair-interactions/s", "cores": cores, "kind": "reference",
           "chain_steps_per_s": value / pairs_per_sweep(N_PART), "chain_steps_per_s_per_core": value / pairs_per_sweep(N_PART) / cores,
           "sample": f"{cores} host cores x 1 chain x {nsweeps} sweeps x {nsteps} steps of oneParticleMoves (N={N_PART}, wall), "
                     f"{total:.1f} s timed inside the workers, reference compiled -O2 -ffp-contract=off"}
    if os.path.exists(ref_so.replace(".so", "_fast.so")):
        vfast, _, _, _ = run_reference(max(2, nsteps // 2), 1, min_step_s=0.5, fast=True)
        out["courtesy_O3_avx2_value"] = vfast
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
    strong_batched = args.scaling == "strong" and args.workload in ("batched", "bulk")
    total_batched = float(Cn) if strong_batched else float(Cn) * world
    if strong_batched:
        shard_b = smcb.shard_chains(int(total_batched), world, rank)
        Cn = shard_b.nchains
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
    mode = smcb.STRICT if args.mode == "strict" else (smcb.FP32 if args.mode == "fp32" else smcb.FAST)
    if mode == smcb.FP32:
        args.kernel = "allparticle"
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
    def droplet_start(n):
        nzl = 4 if n <= 512 else 8
        nx = int(np.ceil(np.sqrt(n / nzl)))
        g = np.array([(i, j, k) for k in range(nzl) for i in range(nx) for j in range(nx)], dtype=float)[:n]
        g[:, 0] = (g[:, 0] - nx / 2) * 1.12
        g[:, 1] = (g[:, 1] - nx / 2) * 1.12
        g[:, 2] = -LZ_BOX / 2 + 0.95 + g[:, 2] * 1.12
        rs = np.random.default_rng(7)
        g += (rs.random(g.shape) * 2 - 1) * 0.05
        return g[rs.permutation(n)].reshape(-1)

    if args.start == "droplet":
        R0 = droplet_start(N)

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
        chain0 = shard_b.chain0 if strong_batched else rank * Cn
    else:
        eng.set_params(smcb.default_params(L=L_BOX, Lz=LZ_BOX, T=TEMP, A=A), GOLDEN_W_M3, ngroups=1)
        chain0 = shard_b.chain0 if strong_batched else rank * Cn
    eng.obs_configure(nebins=64, e_lo=-8.0, e_hi=2.0)
    eng.broadcast_positions(R0)
    eng.set_rng(12345, chain0, 0)
    info = eng.device_info()
    lay = eng.obs_layout()

    # the reduce is done on DELTAS: every step the rank's block (what it gathered since the last reset) is exported,
    # summed over ranks, added to the running job totals below, and reset (Rbin survives the reset)
    # Two scratch blocks alternate so that step k+1's export does not wait for step k's all-reduce, which runs on a side
    # stream UNDER the next sweep launch (ordered after the engine's stream by an event; smcb_obs_export_reset_async).
    obs_cnt = [torch.zeros(lay.u64_total, dtype=torch.int64, device="cuda") for _ in range(2)]
    obs_mom = [torch.zeros(lay.f64_total, dtype=torch.float64, device="cuda") for _ in range(2)]
    tot_cnt = torch.zeros(lay.u64_total, dtype=torch.int64, device="cuda")
    tot_mom = torch.zeros(lay.f64_total, dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()
    ext = torch.cuda.ExternalStream(eng.stream())
    red = {"k": 0, "done": [None, None], "events": [], "last_end": None}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def run_kernel(kernel):
        if kernel == "sweep":
            eng.sweep(S, mode)
        else:
            eng.step_allparticle(S, mode)
        return eng.last_kernel_ms()[0]

    def allreduce_obs():
        """the ONLY collective of the path: the delta the rank gathered since the last reset is exported (and the
        accumulators zeroed) on the engine's stream, summed over ranks with NCCL on a side stream, and added to the
        running job totals there.  Nothing waits for it on the host: it overlaps the next launch."""
        if world == 1:
            return
        k = red["k"]
        red["k"] ^= 1
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(ext)
        if red["done"][k] is not None:
            ext.wait_event(red["done"][k])          # the scratch block is free again (normally long since)
        w1.record(ext)
        eng.obs_export_reset_async(obs_cnt[k].data_ptr(), obs_mom[k].data_ptr())
        ready = torch.cuda.Event(enable_timing=True)
        ready.record(ext)
        side.wait_event(ready)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            t0.record(side)
            smcb.allreduce_observables(obs_cnt[k], obs_mom[k])
            tot_cnt.add_(obs_cnt[k]); tot_mom.add_(obs_mom[k])
            t1.record(side)
        red["done"][k] = t1
        red["events"].append((w0, w1, t0, t1))
        red["last_end"] = ready

    def drain_reduces():
        """wait for the side stream; returns (overlapped all-reduce ms, ms of it that was EXPOSED: the engine's stream
        waiting for a scratch block, plus what remained after the last step's own work)"""
        if world == 1 or not red["events"]:
            red["events"].clear()
            return 0.0, 0.0
        side.synchronize()
        torch.cuda.synchronize()
        total = sum(t0.elapsed_time(t1) for _, _, t0, t1 in red["events"])
        exposed = sum(w0.elapsed_time(w1) for w0, w1, _, _ in red["events"])
        exposed += max(0.0, red["last_end"].elapsed_time(red["events"][-1][3]))
        red["events"].clear()
        return total, exposed

    def one_step(kernel):
        k_ms = run_kernel(kernel)
        pairs = eng.last_pair_counts() + (eng.last_pair_tests(),)
        eng.gather()
        g_ms = eng.last_kernel_ms()[0]
        allreduce_obs()
        return k_ms, g_ms, pairs

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_leg(kernel, nsteps, warmup, sample_clocks=False, want_msd=True):
        """`warmup` untimed steps, then exactly `nsteps` steps between barriers; device times (CUDA events
        on the engine's stream) summed per rank, max over ranks"""
        for _ in range(warmup):
            one_step(kernel)
        drain_reduces()
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        k_tot = g_tot = c_tot = 0.0
        pairs_tot = pairs_cut = pairs_exec = 0
        eng.reset_counters()
        R_before = eng.get_positions() if want_msd else None
        for _ in range(nsteps):
            flush.fill_(1)                      # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize()
            k_ms, g_ms, pairs = one_step(kernel)
            k_tot += k_ms; g_tot += g_ms
            pairs_tot += pairs[0]; pairs_cut += pairs[1]; pairs_exec += pairs[2]
        c_all, c_tot = drain_reduces()           # c_tot: the part of the all-reduces that was NOT hidden under a launch
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        _, na, nt = eng.chain_state()
        dev_ms = smcb.max_over_ranks(k_tot + g_tot + c_tot, device="cuda")
        k_max = smcb.max_over_ranks(k_tot, device="cuda")
        unit_pairs = pairs_per_sweep(N) if kernel == "sweep" else float(N) * (N - 1)
        total_chains = total_batched if args.workload in ("batched", "bulk") else (float(grid[4]) if grid else total_largeN)
        chain_steps = total_chains * S * nsteps
        flops = FLOPS_PAIR * pairs_tot + FLOPS_INCUT * pairs_cut
        flops_exec = FLOPS_PAIR * pairs_exec + FLOPS_INCUT * pairs_cut
        msd = None
        if want_msd:
            # sampling efficiency: mean squared displacement of a molecule over the leg (minimum image in x,y; in z too
            # for the bulk geometry), per second of device time - what an accepted step is worth, for comparing kernels
            dR = (eng.get_positions() - R_before).reshape(Cn, N, 3)
            per = np.array([Lb, Lb, Lb]) if bulk else np.array([L_BOX, L_BOX, np.inf])
            dR -= np.where(np.isfinite(per), per, 1.0) * np.rint(dR / per)
            msd = float(np.mean(np.sum(dR * dR, axis=2)))
        return {"value": chain_steps * unit_pairs / (dev_ms * 1e-3), "chain_steps_per_s": chain_steps / (dev_ms * 1e-3),
                "ms_per_step": dev_ms / nsteps, "kernel_ms_per_step": k_max / nsteps, "gather_ms_per_step": g_tot / nsteps,
                "allreduce_ms_per_step": c_all / nsteps, "allreduce_exposed_ms_per_step": c_tot / nsteps, "wall_s": wall, "clocks": clocks,
                "pairs_in_cutoff_frac": pairs_cut / max(1, pairs_tot), "acceptance": float(na.sum()) / max(1, int(nt.sum())),
                "achieved_tflops": flops / (k_tot * 1e-3) / 1e12, "executed_tflops": flops_exec / (k_tot * 1e-3) / 1e12,
                "pairs_nominal": pairs_tot, "pairs_executed": pairs_exec, "pairs_in_cutoff": pairs_cut, "unit_pairs": unit_pairs,
                "msd_per_leg": msd, "msd_per_s": (msd / (dev_ms * 1e-3)) if msd is not None else None,
                "msd_per_chain_step": (msd / (S * nsteps)) if msd is not None else None}

    fp64_peak, _ = eng.measure_fp64_peak()
    W = max(args.warmup, 3)
    A_tuned = None
    if args.tune_step > 0:
        A_tuned = float(np.median(eng.tune_step_size(args.kernel, mode, target=args.tune_step, rounds=12, nsteps_per_round=max(2, min(S, 20)))))
    main = timed_leg(args.kernel, args.steps, W, sample_clocks=True)

    # ---- end to end through the C-ABI with host buffers (every rank, max over ranks) ----------
    e2e = None
    if not args.no_e2e:
        host_R = torch.empty((Cn, 3 * N), dtype=torch.float64).pin_memory().numpy()
        nsteps_e2e = max(3, min(args.steps, 20))            # the same window of the trajectory as the device-resident leg

        host_E = torch.empty(Cn, dtype=torch.float64).pin_memory().numpy()
        host_na = torch.empty(Cn, dtype=torch.int64).pin_memory().numpy()
        host_nt = torch.empty(Cn, dtype=torch.int64).pin_memory().numpy()

        def e2e_step_separate():
            # SMCB_FP32 is not offered by the pipelined call: the same step through the individual entry points
            eng.set_positions(host_R)
            run_kernel(args.kernel)
            eng.gather()
            allreduce_obs()
            eng.get_positions(host_R)
            return eng.chain_state()

        def e2e_step():
            if mode == smcb.FP32:
                return e2e_step_separate()
            # ONE C-ABI call with host buffers: H2D of the step's positions, energy refresh, S sweeps, gather, D2H of the new
            # positions and the chain state; inside, four chain blocks on four streams overlap their copies and kernels
            eng.sweep_host(host_R, S, mode, kernel=args.kernel, gather=True, E=host_E, naccept=host_na, ntrials=host_nt)
            allreduce_obs()
            return host_E

        # same start as the device-resident leg: the lattice, the same streams, W warm-up steps
        eng.broadcast_positions(R0)
        eng.set_rng(12345, chain0, 0)
        eng.get_positions(host_R)
        for _ in range(W):
            e2e_step()
        drain_reduces()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps_e2e):
            e2e_step()
        drain_reduces()
        barrier()
        te = smcb.max_over_ranks(time.perf_counter() - t0, device="cuda")
        total_chains = total_batched if args.workload in ("batched", "bulk") else (float(grid[4]) if grid else total_largeN)
        e2e = {"value": total_chains * S * nsteps_e2e * main["unit_pairs"] / te, "unit": "pair-interactions/s",
               "chain_steps_per_s": total_chains * S * nsteps_e2e / te, "steps": nsteps_e2e,
               "h2d_bytes_per_step": int(Cn * 3 * N * 8), "d2h_bytes_per_step": int(Cn * 3 * N * 8 + Cn * 24),
               "timing": "host wall clock around smcb_sweep_host (pinned host positions up, energy refresh, sweeps, gather, positions + "
                         "chain state down; four chain blocks pipelined on four streams) + the observable all-reduce; same start "
                         "(lattice + warm-up steps) as the device-resident leg, no L2 flush between steps"}

    # ---- extra legs (reported beside the headline, same JSON line) -----------------------------
    LEG_KEYS = ("value", "chain_steps_per_s", "ms_per_step", "kernel_ms_per_step", "pairs_in_cutoff_frac", "acceptance",
                "msd_per_chain_step", "msd_per_s")
    extra = {}
    if args.workload == "batched" and args.kernel == "sweep" and mode == smcb.FAST and args.start == "lattice" and not args.no_extra:
        short = max(2, min(args.steps, 3))

        def leg(kernel, nsteps, warmup, **more):
            r = timed_leg(kernel, nsteps, warmup)
            out = {k: r[k] for k in LEG_KEYS}
            out.update(roofline_frac_nominal=r["achieved_tflops"] / fp64_peak, roofline_frac_executed=r["executed_tflops"] / fp64_peak, **more)
            return out

        if args.thermalise > 0:
            # sMC's thermalisation (2A, SMC.c:110-125) so molecules reach the wall and partners become common
            eng.set_step_scale(2.0)
            eng.sweep(args.thermalise, mode)
            eng.set_step_scale(1.0)
            extra["thermalised"] = leg("sweep", short, 1, sweeps_before=args.thermalise)
        # north-star kernel B on the same chains: all-particle steps need a small A to be accepted at all; the step size is
        # set per chain by smcb_tune_step_size (target acceptance 0.5) and the sampling efficiency reported as MSD
        eng.set_params(smcb.default_params(L=L_BOX, Lz=LZ_BOX, T=TEMP, A=2e-4), GOLDEN_W_M3, ngroups=1)
        eng.obs_configure(nebins=64, e_lo=-8.0, e_hi=2.0)
        eng.broadcast_positions(R0)
        A_ap = eng.tune_step_size("allparticle", mode, target=0.5, rounds=8, nsteps_per_round=20)
        extra["allparticle_kernel"] = leg("allparticle", short, 2, A_median=float(np.median(A_ap)), A_tuned_for_acceptance=0.5,
                                          note="one all-particle Smart-MC step = N(N-1) ordered pair-interactions; msd_* = mean squared "
                                               "displacement of a molecule per step / per second of device time (sampling efficiency)")
        # the condensed phase (all molecules in a droplet on the wall, the state long reference runs end in): both kernels,
        # step sizes tuned to acceptance 0.5 (sweep) / 0.3 (all-particle)
        Rd = droplet_start(N)
        for kname, target in (("sweep", 0.5), ("allparticle", 0.3)):
            eng.set_params(smcb.default_params(L=L_BOX, Lz=LZ_BOX, T=TEMP, A=0.02 if kname == "sweep" else 1e-5), GOLDEN_W_M3, ngroups=1)
            eng.obs_configure(nebins=64, e_lo=-8.0, e_hi=2.0)
            eng.broadcast_positions(Rd)
            eng.set_rng(12345, chain0, 0)
            if kname == "sweep":
                eng.sweep(100, mode)                 # let the compressed start block relax before tuning
            A_d = eng.tune_step_size(kname, mode, target=target, rounds=8, nsteps_per_round=10 if kname == "sweep" else 40)
            extra[f"droplet_{kname}"] = leg(kname, 2, 1, A_median=float(np.median(A_d)), A_tuned_for_acceptance=target)

    if rank == 0:
        traffic, traffic_note = None, None
        ncu = {}
        npath = os.path.join(ROOT, "profiles", "roofline_ncu.json")
        if os.path.exists(npath):
            ncu = json.load(open(npath)).get(f"{args.kernel}:{args.start}", {})
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
            "scaling": ("strong" if strong_batched else "weak") if args.workload in ("batched", "bulk") else ("weak" if (grid and args.chains) else "strong"),
            "vs_baseline": None,
            "dtype": "f32" if mode == smcb.FP32 else "f64", "data": "synthetic",
            "config": {"workload": (f"{Cn} chains/GPU x N={N} with wall (BASELINE configs[2])" if args.workload == "batched"
                                    else f"{Cn} chains/GPU x N={N} bulk, 3-D periodic, rho*=0.5, T*=1.0 (BASELINE configs[0] geometry)" if bulk
                                    else f"{grid[4]} chains x N={N} on a 16 T x 4 Lz x 4 wall grid x replicas (BASELINE configs[3]), "
                                         f"{Cn} on this rank, {len(grid[1]) * len(grid[2]) * len(grid[3])} observable groups all-reduced"
                                    if grid else f"{int(total_largeN)} chains x N={N} with wall in total (BASELINE configs[4]), {Cn} on this rank")
                                   + f", {args.kernel} kernel, {args.mode}",
                       "chains_per_gpu": Cn, "N": N, "M": M_SITES, "L": Lb if bulk else L_BOX, "Lz": Lb if bulk else LZ_BOX,
                       "T": 1.0 if bulk else TEMP, "A": A if A_tuned is None else A_tuned,
                       "sweeps_per_step": S, "start": ("initializeBox fcc lattice" if args.start == "lattice" else "condensed droplet on the wall") + " + warm-up steps",
                       "l2": "flushed between timed steps (256 MB write)", "rng": "Philox4x32-10",
                       "parallelism": f"chains sharded x{world}, NCCL all-reduce of the observable block only (per step, of the step's delta, "
                                      "on a side stream under the next launch; ms_per_step counts what of it was not hidden)"},
            "kernel_ms_per_step": main["kernel_ms_per_step"], "gather_ms_per_step": main["gather_ms_per_step"],
            "allreduce_ms_per_step": main["allreduce_ms_per_step"], "allreduce_exposed_ms_per_step": main["allreduce_exposed_ms_per_step"],
            "wall_s_timed_region": main["wall_s"],
            "pairs_in_cutoff_frac": main["pairs_in_cutoff_frac"], "acceptance": main["acceptance"],
            "msd_per_chain_step": main["msd_per_chain_step"], "msd_per_s": main["msd_per_s"],
            "roofline": {"bound": "fp64", "achieved": main["achieved_tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": main["achieved_tflops"] / fp64_peak, "frac_nominal": main["achieved_tflops"] / fp64_peak,
                         "frac_executed": main["executed_tflops"] / fp64_peak,
                         "pairs_nominal_per_step": main["pairs_nominal"] / args.steps, "pairs_executed_per_step": main["pairs_executed"] / args.steps,
                         "pairs_in_cutoff_per_step": main["pairs_in_cutoff"] / args.steps,
                         "fp64_pipe_pct": ncu.get("fp64_pipe_pct"), "issue_pct": ncu.get("issue_pct"), "ncu_source": ncu.get("source"),
                         "traffic": traffic, "traffic_note": traffic_note,
                         "note": "frac = frac_nominal: ALGORITHMIC flops (17 per ordered pair-interaction + 16 more inside the cutoff, "
                                 "SURVEY §8d) of a sweep counted as the reference executes it - 2N(N-1) pair-interactions, old + proposed "
                                 "position of every trial - over the measured DFMA peak.  It is a figure of merit, not a utilisation: the "
                                 "kernels execute fewer pair tests (pairs_executed: old-position terms are cached; the all-particle kernel "
                                 "visits every unordered pair once) and run the cutoff test in packed FP32.  frac_executed counts only the "
                                 "pair tests actually executed (17 flops each, FP32) plus 16 FP64 flops per pair inside the cutoff; "
                                 "fp64_pipe_pct / issue_pct are the ncu pipe and issue-slot utilisations of the same kernel on this "
                                 "workload (profiles/, per round).  peak = DFMA peak measured live on this GPU (MEASURED_PEAKS.json has no "
                                 "FP64 entry)"},
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
