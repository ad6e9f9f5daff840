"""ctypes bindings of the parity oracle (TEST INFRASTRUCTURE).

Two things are bound here:
  * `Oracle`  - oracle/liboracle.so, the runtime-sized CPU restatement
                (oracle/smc_oracle.c);
  * `RefLib`  - oracle/_ref/libref_N*_M*.so, the UNMODIFIED reference compiled
                from /root/reference by oracle/build_ref.sh (sizes are macros
                there, so one library per (N, M)).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
RAND_MAX = 2147483647
A0_DEFAULT = 5.960464477539063e-9   # SMC.h:32
B0_DEFAULT = 2.44140625e-5          # SMC.h:33

dptr = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
iptr = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
ulptr = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u8ptr = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class OrcSys(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("L", C.c_double), ("Lz", C.c_double),
                ("rc2", C.c_double), ("a0", C.c_double), ("b0", C.c_double),
                ("periodic_z", C.c_int), ("wall", C.c_int)]


def make_sys(N, M=3, L=33.0, Lz=240.0, rc2=9.0, a0=A0_DEFAULT, b0=B0_DEFAULT, periodic_z=0, wall=1):
    return OrcSys(N, M, L, Lz, rc2, a0, b0, periodic_z, wall)


class Oracle:
    def __init__(self, fast=False):
        name = "liboracle_fast.so" if fast else "liboracle.so"
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(path)
        S = C.POINTER(OrcSys)
        L.orc_sizeof_sys.restype = C.c_size_t
        assert L.orc_sizeof_sys() == C.sizeof(OrcSys)
        L.orc_energy_single.restype = C.c_double
        L.orc_energy_single.argtypes = [S, dptr, C.c_int]
        L.orc_force_single.argtypes = [S, dptr, C.c_int, dptr]
        L.orc_energy.restype = C.c_double
        L.orc_energy.argtypes = [S, dptr]
        L.orc_forces.argtypes = [S, dptr, dptr]
        L.orc_pressure.restype = C.c_double
        L.orc_pressure.argtypes = [S, dptr]
        L.orc_walls_energy_single.restype = C.c_double
        L.orc_walls_energy_single.argtypes = [S, C.c_double, C.c_double, C.c_double, dptr]
        L.orc_walls_force.argtypes = [S, C.c_double, C.c_double, C.c_double, dptr, dptr]
        L.orc_walls_energy.restype = C.c_double
        L.orc_walls_energy.argtypes = [S, dptr, dptr]
        L.orc_walls_pressure.restype = C.c_double
        L.orc_walls_pressure.argtypes = [S, dptr, dptr]
        L.orc_walls_virial_intended.restype = C.c_double
        L.orc_walls_virial_intended.argtypes = [S, dptr, dptr]
        L.orc_box_muller.argtypes = [C.c_double, C.c_size_t, iptr, C.c_int, dptr]
        L.orc_sweep.argtypes = [S, dptr, dptr, dptr, C.c_double, C.c_double, dptr, C.c_longlong, dptr,
                                C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p]
        L.orc_sweep_from_ints.argtypes = [S, dptr, dptr, dptr, C.c_double, C.c_double, iptr, C.c_int,
                                          C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_expand_stream.argtypes = [C.c_int, C.c_double, iptr, C.c_int, dptr, C.POINTER(C.c_longlong), dptr]
        L.orc_local_density.argtypes = [S, dptr, C.c_int, C.c_int, ulptr, iptr, ulptr]
        L.orc_bounds_check.restype = C.c_int
        L.orc_bounds_check.argtypes = [S, dptr, C.c_double, C.POINTER(C.c_int)]
        L.orc_initialize_box.restype = C.c_int
        L.orc_initialize_box.argtypes = [C.c_double, C.c_double, C.c_int, dptr]
        L.orc_fcc_lattice.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dptr]
        L.orc_walls_from_gauss.argtypes = [C.c_int, C.c_double, C.c_double, dptr, dptr, dptr]
        u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_rng_particle.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, dptr, C.POINTER(C.c_double)]
        L.orc_rng_step_scalars.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.orc_total.argtypes = [S, dptr, dptr, dptr, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_allparticle_step.restype = C.c_int
        L.orc_allparticle_step.argtypes = [S, dptr, dptr, C.POINTER(C.c_double), dptr, C.c_double, C.c_double,
                                           dptr, C.c_double, C.POINTER(C.c_double)]
        L.orc_run_sweeps.restype = C.c_long
        L.orc_run_sweeps.argtypes = [S, dptr, dptr, C.c_double, C.c_double, C.c_int, C.c_uint, C.POINTER(C.c_double)]

    # --- thin pythonic wrappers -------------------------------------------------
    def energy_single(self, s, r, i):
        return self.lib.orc_energy_single(C.byref(s), r, i)

    def force_single(self, s, r, i):
        F = np.zeros(3)
        self.lib.orc_force_single(C.byref(s), r, i, F)
        return F

    def energy(self, s, r):
        return self.lib.orc_energy(C.byref(s), r)

    def forces(self, s, r, F=None):
        F = np.zeros(3 * s.N) if F is None else F
        self.lib.orc_forces(C.byref(s), r, F)
        return F

    def pressure(self, s, r):
        return self.lib.orc_pressure(C.byref(s), r)

    def walls_energy_single(self, s, p, W):
        return self.lib.orc_walls_energy_single(C.byref(s), p[0], p[1], p[2], W)

    def walls_force(self, s, p, W, F=None):
        F = np.zeros(3) if F is None else F
        self.lib.orc_walls_force(C.byref(s), p[0], p[1], p[2], W, F)
        return F

    def walls_energy(self, s, r, W):
        return self.lib.orc_walls_energy(C.byref(s), r, W)

    def walls_pressure(self, s, r, W):
        return self.lib.orc_walls_pressure(C.byref(s), r, W)

    def walls_virial_intended(self, s, r, W):
        return self.lib.orc_walls_virial_intended(C.byref(s), r, W)

    def box_muller(self, sigma, length, rnd):
        out = np.zeros(length)
        self.lib.orc_box_muller(sigma, length, np.ascontiguousarray(rnd, dtype=np.int32), RAND_MAX, out)
        return out

    def expand_stream(self, N, A, rnd):
        displ = np.zeros(3 * N)
        u = np.zeros(N)
        off = C.c_longlong(0)
        self.lib.orc_expand_stream(N, A, np.ascontiguousarray(rnd, dtype=np.int32), RAND_MAX, displ, C.byref(off), u)
        return displ, off.value, u

    def sweep(self, s, R, W, A, T, displ, offset, u, E=0.0, want_flags=False):
        """returns (naccept, E_after[, flags]); R updated in place"""
        Rn = np.zeros_like(R)
        j = C.c_int(0)
        e = C.c_double(E)
        flags = np.zeros(s.N, dtype=np.uint8) if want_flags else None
        self.lib.orc_sweep(C.byref(s), R, Rn, W, A, T, displ, int(offset), u, C.byref(j), C.byref(e),
                           flags.ctypes.data if want_flags else None)
        return (j.value, e.value, flags) if want_flags else (j.value, e.value)

    def sweep_from_ints(self, s, R, W, A, T, rnd, E=0.0):
        Rn = np.zeros_like(R)
        j = C.c_int(0)
        e = C.c_double(E)
        self.lib.orc_sweep_from_ints(C.byref(s), R, Rn, W, A, T, np.ascontiguousarray(rnd, dtype=np.int32),
                                     RAND_MAX, C.byref(j), C.byref(e))
        return j.value, e.value

    def local_density(self, s, r, D, Rbin, Mu, ncx=33, ncz=33):
        self.lib.orc_local_density(C.byref(s), r, ncx, ncz, D, Rbin, Mu)

    def initialize_box(self, L, Lz, n):
        X = np.zeros(3 * n)
        sites = self.lib.orc_initialize_box(L, Lz, n, X)
        return X, sites

    def fcc_lattice(self, L, Lz, nx, ny, nz):
        X = np.zeros(3 * 4 * nx * ny * nz)
        self.lib.orc_fcc_lattice(L, Lz, nx, ny, nz, X)
        return X

    def total(self, s, r, W):
        F = np.zeros(3 * s.N)
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self.lib.orc_total(C.byref(s), r, W, F, C.byref(a), C.byref(b), C.byref(c))
        return F, a.value, b.value, c.value

    def allparticle_step(self, s, R, F, U, W, A, T, xi, u):
        """returns (accepted, U_after, ln_ap); R, F updated in place on acceptance"""
        uu = C.c_double(U)
        ln = C.c_double(0)
        ok = self.lib.orc_allparticle_step(C.byref(s), R, F, C.byref(uu), W, A, T, xi, u, C.byref(ln))
        return ok, uu.value, ln.value

    def philox(self, ctr, key):
        out = np.zeros(4, dtype=np.uint32)
        self.lib.orc_philox4x32_10(np.asarray(ctr, dtype=np.uint32), np.asarray(key, dtype=np.uint32), out)
        return out

    def rng_particle(self, seed, chain, step, particle):
        g = np.zeros(3)
        u = C.c_double(0)
        self.lib.orc_rng_particle(seed, chain, step, particle, g, C.byref(u))
        return g, u.value

    def rng_step_scalars(self, seed, chain, step):
        off = C.c_uint32(0)
        u = C.c_double(0)
        self.lib.orc_rng_step_scalars(seed, chain, step, C.byref(off), C.byref(u))
        return off.value, u.value

    def run_sweeps(self, s, R, W, A, T, nsweeps, seed=1, E=0.0):
        e = C.c_double(E)
        acc = self.lib.orc_run_sweeps(C.byref(s), R, W, A, T, nsweeps, seed, C.byref(e))
        return acc, e.value


class RefLib:
    """The compiled reference for one (N, M).  Functions keep the reference's names."""

    def __init__(self, N, M=3, fast=False, path=None):
        """path: another library exporting the same SMC.h API (the drop-in, tests/dropin/_build)"""
        if path is None:
            tag = f"libref_N{N}_M{M}" + ("_fast" if fast else "") + ".so"
            path = os.path.join(ORACLE_DIR, "_ref", tag)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle/build_ref.sh where /root/reference exists "
                                    "(or make -C tests/dropin for the drop-in)")
        self.N, self.M = N, M
        L = self.lib = C.CDLL(path)
        assert L.ref_N() == N and L.ref_M() == M
        L.energySingle.restype = C.c_double
        L.energySingle.argtypes = [dptr, C.c_double, C.c_int]
        L.forceSingle.argtypes = [dptr, C.c_double, C.c_int] + [C.POINTER(C.c_double)] * 3
        L.energy.restype = C.c_double
        L.energy.argtypes = [dptr, C.c_double]
        L.forces.argtypes = [dptr, C.c_double, dptr]
        L.pressure.restype = C.c_double
        L.pressure.argtypes = [dptr, C.c_double, C.c_double]
        L.wallsEnergySingle.restype = C.c_double
        L.wallsEnergySingle.argtypes = [C.c_double] * 3 + [dptr, C.c_double, C.c_double]
        L.wallsForce.argtypes = [C.c_double] * 3 + [dptr, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 3
        L.wallsEnergy.restype = C.c_double
        L.wallsEnergy.argtypes = [dptr, dptr, C.c_double, C.c_double]
        L.wallsPressure.restype = C.c_double
        L.wallsPressure.argtypes = [dptr, dptr, C.c_double, C.c_double]
        L.vecBoxMuller.argtypes = [C.c_double, C.c_size_t, dptr]
        L.oneParticleMoves.argtypes = [dptr, dptr, dptr] + [C.c_double] * 4 + [C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.localDensityAndMobility.argtypes = [dptr, C.c_double, C.c_double, ulptr, iptr, ulptr]
        L.initializeBox.argtypes = [C.c_double, C.c_double, C.c_int, dptr]
        L.boundsCheck.restype = C.c_int
        L.boundsCheck.argtypes = [dptr, C.c_double, C.c_double]
        L.oracle_set_replay.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_replay_pos.restype = C.c_size_t
        L.oracle_replay_underflow.restype = C.c_size_t
        L.oracle_srand.argtypes = [C.c_uint]
        L.ref_sizeof_sim.restype = C.c_size_t
        self._replay = None

    def set_replay(self, rnd):
        if rnd is None:
            self._replay = None
            self.lib.oracle_set_replay(None, 0)
        else:
            self._replay = np.ascontiguousarray(rnd, dtype=np.int32)  # keep alive
            self.lib.oracle_set_replay(self._replay.ctypes.data, self._replay.size)

    def replay_pos(self):
        return self.lib.oracle_replay_pos()

    def replay_underflow(self):
        return self.lib.oracle_replay_underflow()

    def energySingle(self, r, L, i):
        return self.lib.energySingle(r, L, i)

    def forceSingle(self, r, L, i):
        fx, fy, fz = C.c_double(0), C.c_double(0), C.c_double(0)
        self.lib.forceSingle(r, L, i, C.byref(fx), C.byref(fy), C.byref(fz))
        return np.array([fx.value, fy.value, fz.value])

    def energy(self, r, L):
        return self.lib.energy(r, L)

    def forces(self, r, L, F=None):
        F = np.zeros(3 * self.N) if F is None else F
        self.lib.forces(r, L, F)
        return F

    def pressure(self, r, L, Lz):
        return self.lib.pressure(r, L, Lz)

    def wallsEnergySingle(self, p, W, L, Lz):
        return self.lib.wallsEnergySingle(p[0], p[1], p[2], W, L, Lz)

    def wallsForce(self, p, W, L, Lz, F0=(0.0, 0.0, 0.0)):
        fx, fy, fz = C.c_double(F0[0]), C.c_double(F0[1]), C.c_double(F0[2])
        self.lib.wallsForce(p[0], p[1], p[2], W, L, Lz, C.byref(fx), C.byref(fy), C.byref(fz))
        return np.array([fx.value, fy.value, fz.value])

    def wallsEnergy(self, r, W, L, Lz):
        return self.lib.wallsEnergy(r, W, L, Lz)

    def wallsPressure(self, r, W, L, Lz):
        return self.lib.wallsPressure(r, W, L, Lz)

    def vecBoxMuller(self, sigma, length):
        out = np.zeros(length)
        self.lib.vecBoxMuller(sigma, length, out)
        return out

    def oneParticleMoves(self, R, W, L, Lz, A, T, E=0.0):
        Rn = np.zeros_like(R)
        j = C.c_int(0)
        e = C.c_double(E)
        self.lib.oneParticleMoves(R, Rn, W, L, Lz, A, T, C.byref(j), C.byref(e))
        return j.value, e.value

    def localDensityAndMobility(self, r, L, Lz, D, Rbin, Mu):
        self.lib.localDensityAndMobility(r, L, Lz, D, Rbin, Mu)

    def initializeBox(self, L, Lz):
        X = np.zeros(3 * self.N)
        self.lib.initializeBox(L, Lz, self.N, X)
        return X


class RefNoWall:
    """Bulk prototype (SMC_noMPI_noWall.c): energy/forces/pressure only."""

    def __init__(self, N):
        path = os.path.join(ORACLE_DIR, "_ref", f"libref_nowall_N{N}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.N = N
        L = self.lib = C.CDLL(path)
        assert L.ref_N() == N
        L.energy.restype = C.c_double
        L.energy.argtypes = [dptr, C.c_double]
        L.forces.argtypes = [dptr, C.c_double, dptr]
        L.pressure.restype = C.c_double
        L.pressure.argtypes = [dptr, C.c_double]
        L.initializeBox.argtypes = [C.c_double, C.c_int, dptr]

    def energy(self, r, L):
        return self.lib.energy(r, L)

    def forces(self, r, L):
        F = np.zeros(3 * self.N)
        self.lib.forces(r, L, F)
        return F

    def pressure(self, r, L):
        return self.lib.pressure(r, L)

    def initializeBox(self, L):
        X = np.zeros(3 * self.N)
        self.lib.initializeBox(L, self.N, X)
        return X


# Survey-generated wall parameters: initializeWalls(1.6,0,3.0,0.5) after srand(42), M=3
GOLDEN_W_M3 = np.array([
    962.2264072645321, 57.35316319850277,
    874.39446992695275, 52.11797177356199,
    857.36680597299653, 51.103043912231705,
    1024.1964124687327, 61.046863345428257,
    925.40789594507817, 55.158608910148025,
    913.63518965684239, 54.456900933792724,
    848.90539177252572, 50.598704324515197,
    992.35137245273086, 59.148751047416368,
    844.42493013196849, 50.331648000000015,
])


def random_walls(M, rng):
    """W[2m], W[2m+1] in the range initializeWalls produces (x0=1.6, ymin~N(3,0.5))."""
    ymin = 3.0 + 0.5 * rng.standard_normal(M * M)
    W = np.empty(2 * M * M)
    W[0::2] = 1.6 ** 12 * ymin
    W[1::2] = 1.6 ** 6 * ymin
    return W


def config_gas(N, L, Lz, rng, zfrac=0.45):
    """uniform random gas inside the slab"""
    R = np.empty(3 * N)
    R[0::3] = (rng.random(N) - 0.5) * L
    R[1::3] = (rng.random(N) - 0.5) * L
    R[2::3] = (rng.random(N) * 2 - 1) * zfrac * Lz
    return R


def config_slab(N, L, Lz, rng, thickness=6.0, dmin=0.92):
    """layer adsorbed on the lower wall (z from -Lz/2+0.8) by random sequential insertion with
    minimum pair distance dmin: a moderate number of in-cutoff pairs, some of them close"""
    pts = np.empty((N, 3))
    n = 0
    tries = 0
    while n < N:
        tries += 1
        assert tries < 5000000
        p = np.array([(rng.random() - 0.5) * L, (rng.random() - 0.5) * L,
                      -Lz / 2 + 0.8 + rng.random() * thickness])
        if n:
            d = pts[:n] - p
            d[:, 0] -= L * np.rint(d[:, 0] / L)
            d[:, 1] -= L * np.rint(d[:, 1] / L)
            if np.min(np.einsum("ij,ij->i", d, d)) < dmin * dmin:
                continue
        pts[n] = p
        n += 1
    return pts.reshape(-1).copy()


def config_droplet(N, L, Lz, rng, spacing=1.12, jitter=0.07, nz=4, z0=0.95):
    """dense droplet sitting on the lower wall: jittered simple-cubic block (liquid-like density,
    ~80 neighbours inside the cutoff, pair terms up to ~1e2..1e3) - the condensed state of SMC.c runs"""
    nxy = int(np.ceil(np.sqrt(N / nz)))
    assert nxy * spacing < L
    g = np.array([(i, j, k) for k in range(nz) for i in range(nxy) for j in range(nxy)], dtype=float)[:N]
    g[:, 0] = (g[:, 0] - nxy / 2) * spacing
    g[:, 1] = (g[:, 1] - nxy / 2) * spacing
    g[:, 2] = -Lz / 2 + z0 + g[:, 2] * spacing
    g += (rng.random(g.shape) * 2 - 1) * jitter
    perm = rng.permutation(N)
    return g[perm].reshape(-1).copy()
