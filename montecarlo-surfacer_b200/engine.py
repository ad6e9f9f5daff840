"""ctypes binding of include/smcb200.h (one Engine = one smcb_engine handle = one GPU).

Method names follow the C entry points; array arguments are numpy arrays in the reference's
layouts (positions AoS [chain][3N], SMC.h:84; W interleaved (a,b), SMC.c:745-760).
"""
import ctypes as C
import os
import re

import numpy as np

FAST, STRICT, FP32 = 0, 1, 2
WALL, PERIODIC_Z = 1, 2
A0_DEFAULT = 5.960464477539063e-9   # SMC.h:32
B0_DEFAULT = 2.44140625e-5          # SMC.h:33

# The wall table main.c builds: initializeWalls(1.6, 0.0, 3.0, 0.5, W, f) after its srand(42), M = 3 (main.c:74-87,
# SMC.c:475-501; values from SURVEY.md App. D, glibc rand).  W[2m] = a, W[2m+1] = b, m = i*M + j.
REFERENCE_WALL_M3 = np.array([
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

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


class SmcbError(RuntimeError):
    pass


class ChainParams(C.Structure):
    """smcb_chain_params"""
    _fields_ = [("L", C.c_double), ("Lz", C.c_double), ("T", C.c_double), ("A", C.c_double),
                ("rc2", C.c_double), ("zwall_a", C.c_double), ("zwall_b", C.c_double),
                ("flags", C.c_uint32), ("wall", C.c_uint32), ("group", C.c_uint32), ("pad_", C.c_uint32)]


class ObsLayout(C.Structure):
    """smcb_obs_layout"""
    _fields_ = [("ngroups", C.c_int), ("nvox", C.c_int), ("nz", C.c_int), ("nebins", C.c_int),
                ("e_lo", C.c_double), ("e_hi", C.c_double),
                ("u64_per_group", C.c_size_t), ("f64_per_group", C.c_size_t),
                ("u64_total", C.c_size_t), ("f64_total", C.c_size_t)]


def default_params(L=33.0, Lz=240.0, T=1.1, A=None, rc2=9.0, a0=A0_DEFAULT, b0=B0_DEFAULT,
                   flags=WALL, wall=0, group=0):
    """main.c geometry (main.c:35-51): L=33, Lz=240 (N>=150), A = gamma*T with gamma=1"""
    return ChainParams(L, Lz, T, T if A is None else A, rc2, a0, b0, flags, wall, group, 0)


def lib_path():
    """in-tree libsmcb200.so; SMCB200_LIB selects another build of the SAME library (kernel experiments)"""
    return os.environ.get("SMCB200_LIB") or os.path.join(_HERE, "libsmcb200.so")


_lib = None


def load_library():
    """Load libsmcb200.so or fail loudly - there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SmcbError(f"{path} not found: build it with `make -C montecarlo-surfacer_b200/csrc` "
                        "(or __graft_entry__.build()); smcb200 has no CPU fallback")
    lib = C.CDLL(path)
    P = C.c_void_p
    dp = C.POINTER(C.c_double)
    sig = {
        "smcb_create": [C.POINTER(P), C.c_int, C.c_int, C.c_int, C.c_int],
        "smcb_destroy": [P],
        "smcb_device_info": [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)],
        "smcb_set_params": [P, C.POINTER(ChainParams), C.c_int, P, C.c_int, C.c_int],
        "smcb_set_positions": [P, P],
        "smcb_get_positions": [P, P],
        "smcb_broadcast_positions": [P, P],
        "smcb_set_rng": [P, C.c_uint64, C.c_uint32, C.c_uint64],
        "smcb_evaluate": [P, C.c_int] + [P] * 8,
        "smcb_get_wall_virial": [P, P],
        "smcb_obs_set_wall_virial": [P, C.c_int],
        "smcb_sweep_fed": [P, C.c_int, C.c_int, P, P, P, P],
        "smcb_sweep": [P, C.c_int, C.c_int],
        "smcb_sweep_traced": [P, C.c_int, C.c_int, P, P, P, P, P],
        "smcb_set_rbin": [P, P],
        "smcb_set_step_scale": [P, C.c_double],
        "smcb_step_allparticle_fed": [P, C.c_int, C.c_int, P, P, P, P],
        "smcb_step_allparticle": [P, C.c_int, C.c_int],
        "smcb_refresh_energy": [P, C.c_int],
        "smcb_get_chain_state": [P, P, P, P],
        "smcb_set_chain_energy": [P, P],
        "smcb_reset_counters": [P],
        "smcb_obs_configure": [P, C.c_int, C.c_double, C.c_double],
        "smcb_obs_layout_get": [P, C.POINTER(ObsLayout)],
        "smcb_gather": [P],
        "smcb_obs_reset": [P],
        "smcb_obs_get": [P, P, P],
        "smcb_obs_export_device": [P, P, P],
        "smcb_obs_import_device": [P, P, P],
        "smcb_get_rbin": [P, P],
        "smcb_checkpoint_save": [P, C.c_char_p],
        "smcb_checkpoint_load": [P, C.c_char_p],
        "smcb_last_kernel_ms": [P, C.POINTER(C.c_float), C.POINTER(C.c_int)],
        "smcb_last_pair_counts": [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
        "smcb_last_pair_tests": [P, C.POINTER(C.c_uint64)],
        "smcb_debug_sweep_stats": [P, P],
        "smcb_sweep_host": [P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, P, P],
        "smcb_obs_allreduce_teardown": [],
        "smcb_tune_step_size": [P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int],
        "smcb_obs_export_reset_async": [P, P, P],
        "smcb_get_step_sizes": [P, P],
        "smcb_measure_fp64_peak": [P, dp, C.POINTER(C.c_float)],
        "smcb_device_positions": [P, C.POINTER(P), C.POINTER(C.c_size_t), C.POINTER(C.c_int)],
        "smcb_debug_capture_cache": [P, C.c_int],
        "smcb_debug_get_cache": [P, P, P, P],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.smcb_last_error.restype = C.c_char_p
    lib.smcb_stream.argtypes = [P]
    lib.smcb_stream.restype = P
    _lib = lib
    return lib


def header_symbols():
    """every function name declared in include/smcb200.h"""
    text = open(os.path.join(_ROOT, "include", "smcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb_[a-z0-9_]+)\s*\(", text)))


def exported_symbols():
    """dynamic symbols libsmcb200.so exports (via ctypes lookup of the header's names)"""
    lib = load_library()
    return [s for s in header_symbols() if hasattr(lib, s)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


class Engine:
    """A batch of `nchains` independent chains of `N` particles on one GPU."""

    def __init__(self, nchains, N, M=3, device=0):
        self.lib = load_library()
        self.C, self.N, self.M = int(nchains), int(N), int(M)
        self._h = C.c_void_p()
        rc = self.lib.smcb_create(C.byref(self._h), device, self.C, self.N, self.M)
        if rc != 0:
            self._h = None
            raise SmcbError(f"smcb_create failed ({rc}): {self.lib.smcb_last_error().decode()}")

    # -- plumbing ------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise SmcbError(f"smcb error {rc}: {self.lib.smcb_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.smcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._ck(self.lib.smcb_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": mem.value}

    # -- inputs --------------------------------------------------------------
    def set_params(self, params, W=None, ngroups=1):
        """params: one ChainParams or a sequence of nchains; W: [nwalls][2*M*M]"""
        if isinstance(params, ChainParams):
            arr = (ChainParams * 1)(params)
            n = 1
        else:
            params = list(params)
            n = len(params)
            arr = (ChainParams * n)(*params)
        if W is None:
            Wc, nw = None, 0
        else:
            Wc = _f64(W)
            if Wc.size % (2 * self.M * self.M):
                raise ValueError("W must hold whole tables of 2*M*M doubles")
            nw = Wc.size // (2 * self.M * self.M)
        self._ck(self.lib.smcb_set_params(self._h, arr, n, _ptr(Wc), nw, ngroups))

    def set_positions(self, R):
        R = _f64(R, self.C * 3 * self.N)
        self._ck(self.lib.smcb_set_positions(self._h, _ptr(R)))

    def broadcast_positions(self, R0):
        R0 = _f64(R0, 3 * self.N)
        self._ck(self.lib.smcb_broadcast_positions(self._h, _ptr(R0)))

    def get_positions(self, out=None):
        R = np.empty((self.C, 3 * self.N)) if out is None else out
        self._ck(self.lib.smcb_get_positions(self._h, _ptr(R)))
        return R

    def set_rng(self, seed, chain0=0, step0=0):
        self._ck(self.lib.smcb_set_rng(self._h, seed, chain0, step0))

    def set_step_scale(self, scale):
        self._ck(self.lib.smcb_set_step_scale(self._h, scale))

    # -- static evaluation -----------------------------------------------------
    def evaluate(self, mode=FAST, per_particle=True):
        """dict with e_lj[C,N], f_lj[C,3N], e_wall[C,N], f_wall[C,3N] (if per_particle) and the
        chain totals U_lj, U_wall, vir_lj, vir_wall_ref [C]"""
        Cn, N = self.C, self.N
        out = {k: np.empty(Cn) for k in ("U_lj", "U_wall", "vir_lj", "vir_wall_ref")}
        if per_particle:
            out.update(e_lj=np.empty((Cn, N)), f_lj=np.empty((Cn, 3 * N)),
                       e_wall=np.empty((Cn, N)), f_wall=np.empty((Cn, 3 * N)))
        g = out.get
        self._ck(self.lib.smcb_evaluate(self._h, mode, _ptr(g("e_lj")), _ptr(g("f_lj")), _ptr(g("e_wall")),
                                        _ptr(g("f_wall")), _ptr(out["U_lj"]), _ptr(out["U_wall"]),
                                        _ptr(out["vir_lj"]), _ptr(out["vir_wall_ref"])))
        return out

    def wall_virial(self):
        """chain sums of the wall virial as the reference meant it (after evaluate() or gather())"""
        v = np.empty(self.C)
        self._ck(self.lib.smcb_get_wall_virial(self._h, _ptr(v)))
        return v

    def obs_set_wall_virial(self, intended):
        self._ck(self.lib.smcb_obs_set_wall_virial(self._h, int(bool(intended))))

    # -- the sweep ---------------------------------------------------------------
    def sweep_fed(self, displ, offset, u, mode=STRICT, want_accepted=False):
        """displ [S,C,3N], offset [S,C] (int64), u [S,C,N]"""
        displ = _f64(displ)
        S = displ.size // (self.C * 3 * self.N)
        offset = np.ascontiguousarray(offset, dtype=np.int64)
        u = _f64(u, S * self.C * self.N)
        if displ.size != S * self.C * 3 * self.N or offset.size != S * self.C:
            raise ValueError("fed sweep input shapes do not agree")
        acc = np.zeros((S, self.C, self.N), dtype=np.uint8) if want_accepted else None
        self._ck(self.lib.smcb_sweep_fed(self._h, S, mode, _ptr(displ), _ptr(offset), _ptr(u), _ptr(acc)))
        return acc

    def sweep(self, nsweeps, mode=FAST):
        self._ck(self.lib.smcb_sweep(self._h, nsweeps, mode))

    def sweep_host(self, R, nsteps, mode=FAST, kernel="sweep", gather=False, E=None, naccept=None, ntrials=None):
        """smcb_sweep_host: R (host, [C, 3N] float64, C-contiguous, ideally pinned) is uploaded, advanced nsteps sweeps
        (kernel="sweep") or all-particle steps ("allparticle"), optionally gathered, and overwritten IN PLACE with the
        new positions; E / naccept / ntrials (optional preallocated arrays) receive the chain state.  Copies and
        kernels of the four chain blocks overlap."""
        if not (isinstance(R, np.ndarray) and R.dtype == np.float64 and R.flags.c_contiguous and R.size == self.C * 3 * self.N):
            raise ValueError("R must be a C-contiguous float64 array of nchains*3N elements (it is updated in place)")
        for arr, dt in ((E, np.float64), (naccept, np.int64), (ntrials, np.int64)):
            if arr is not None and not (arr.dtype == dt and arr.flags.c_contiguous and arr.size == self.C):
                raise ValueError("chain-state outputs must be contiguous arrays of nchains elements")
        self._ck(self.lib.smcb_sweep_host(self._h, _ptr(R), nsteps, mode, 0 if kernel == "sweep" else 1, 1 if gather else 0,
                                          _ptr(E), _ptr(naccept), _ptr(ntrials)))

    def sweep_traced(self, nsweeps, mode=FAST, displ=None, offset=None, u=None):
        """(E_trace [S,C], acc_trace [S,C]): sMC's E[n+1] and jj[n] of every sweep"""
        Et = np.empty((nsweeps, self.C))
        at = np.empty((nsweeps, self.C), dtype=np.int32)
        if displ is not None:
            displ, u = _f64(displ, nsweeps * self.C * 3 * self.N), _f64(u, nsweeps * self.C * self.N)
            offset = np.ascontiguousarray(offset, dtype=np.int64)
        self._ck(self.lib.smcb_sweep_traced(self._h, nsweeps, mode, _ptr(displ), _ptr(offset), _ptr(u), _ptr(Et), _ptr(at)))
        return Et, at

    # -- the all-particle step -----------------------------------------------------
    def step_allparticle_fed(self, xi, u, mode=FAST):
        """xi [S,C,3N] (already scaled by sqrt(2A)), u [S,C]; returns (lnap [S,C], accepted [S,C])"""
        xi = _f64(xi)
        S = xi.size // (self.C * 3 * self.N)
        u = _f64(u, S * self.C)
        lnap = np.empty((S, self.C))
        acc = np.zeros((S, self.C), dtype=np.uint8)
        self._ck(self.lib.smcb_step_allparticle_fed(self._h, S, mode, _ptr(xi), _ptr(u), _ptr(lnap), _ptr(acc)))
        return lnap, acc

    def step_allparticle(self, nsteps, mode=FAST):
        self._ck(self.lib.smcb_step_allparticle(self._h, nsteps, mode))

    def obs_export_reset_async(self, counters_ptr, moments_ptr):
        """device pointers (e.g. torch tensors' data_ptr()); asynchronous on the engine's stream (Engine.stream())"""
        self._ck(self.lib.smcb_obs_export_reset_async(self._h, C.c_void_p(counters_ptr), C.c_void_p(moments_ptr)))

    def tune_step_size(self, kernel="sweep", mode=FAST, target=0.5, rounds=8, nsteps_per_round=10):
        """smcb_tune_step_size: per-chain A adapted towards the target acceptance; returns the A array"""
        self._ck(self.lib.smcb_tune_step_size(self._h, 0 if kernel == "sweep" else 1, mode, target, rounds, nsteps_per_round))
        A = np.empty(self.C)
        self._ck(self.lib.smcb_get_step_sizes(self._h, _ptr(A)))
        return A

    # -- chain state -------------------------------------------------------------
    def refresh_energy(self, mode=FAST):
        self._ck(self.lib.smcb_refresh_energy(self._h, mode))

    def chain_state(self):
        E = np.empty(self.C)
        na = np.empty(self.C, dtype=np.int64)
        nt = np.empty(self.C, dtype=np.int64)
        self._ck(self.lib.smcb_get_chain_state(self._h, _ptr(E), _ptr(na), _ptr(nt)))
        return E, na, nt

    def set_chain_energy(self, E):
        E = _f64(E, self.C)
        self._ck(self.lib.smcb_set_chain_energy(self._h, _ptr(E)))

    def reset_counters(self):
        self._ck(self.lib.smcb_reset_counters(self._h))

    # -- observables ---------------------------------------------------------------
    def obs_configure(self, nebins=64, e_lo=-8.0, e_hi=2.0):
        self._ck(self.lib.smcb_obs_configure(self._h, nebins, e_lo, e_hi))

    def obs_layout(self):
        lay = ObsLayout()
        self._ck(self.lib.smcb_obs_layout_get(self._h, C.byref(lay)))
        return lay

    def gather(self):
        self._ck(self.lib.smcb_gather(self._h))

    def obs_reset(self):
        self._ck(self.lib.smcb_obs_reset(self._h))

    def obs_get(self):
        """returns dict per group index: D, Mu [33,33,33], zprof [33], ehist, nsamples, and moments"""
        lay = self.obs_layout()
        cnt = np.zeros(lay.u64_total, dtype=np.uint64)
        mom = np.zeros(lay.f64_total)
        self._ck(self.lib.smcb_obs_get(self._h, _ptr(cnt), _ptr(mom)))
        return unpack_obs(lay, cnt, mom)

    def obs_export_device(self, counters_ptr, moments_ptr):
        self._ck(self.lib.smcb_obs_export_device(self._h, counters_ptr, moments_ptr))

    def obs_import_device(self, counters_ptr, moments_ptr):
        self._ck(self.lib.smcb_obs_import_device(self._h, counters_ptr, moments_ptr))

    def rbin(self):
        rb = np.empty((self.C, self.N), dtype=np.int32)
        self._ck(self.lib.smcb_get_rbin(self._h, _ptr(rb)))
        return rb

    def set_rbin(self, rb):
        rb = np.ascontiguousarray(rb, dtype=np.int32)
        self._ck(self.lib.smcb_set_rbin(self._h, _ptr(rb)))

    # -- checkpoint / resume ------------------------------------------------------------
    def checkpoint_save(self, path):
        self._ck(self.lib.smcb_checkpoint_save(self._h, os.fsencode(path)))

    def checkpoint_load(self, path):
        self._ck(self.lib.smcb_checkpoint_load(self._h, os.fsencode(path)))

    # -- measurement -----------------------------------------------------------------
    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self.lib.smcb_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_pair_counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.smcb_last_pair_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_pair_tests(self):
        a = C.c_uint64()
        self._ck(self.lib.smcb_last_pair_tests(self._h, C.byref(a)))
        return a.value

    def debug_sweep_stats(self):
        a = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.smcb_debug_sweep_stats(self._h, _ptr(a)))
        return a

    def measure_fp64_peak(self):
        t, ms = C.c_double(), C.c_float()
        self._ck(self.lib.smcb_measure_fp64_peak(self._h, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def debug_capture_cache(self, on=True):
        self._ck(self.lib.smcb_debug_capture_cache(self._h, int(on)))

    def debug_get_cache(self):
        """(e_tot [C,N], f_tot [C,3N], nb [C,N]) held by the FAST sweep kernel at its end"""
        e = np.empty((self.C, self.N))
        f = np.empty((self.C, 3 * self.N))
        nb = np.empty((self.C, self.N))
        self._ck(self.lib.smcb_debug_get_cache(self._h, _ptr(e), _ptr(f), _ptr(nb)))
        return e, f, nb

    def stream(self):
        return self.lib.smcb_stream(self._h)


def obs_allreduce(engines):
    """smcb_obs_allreduce: sum the observable blocks of several engines of THIS process (one per GPU) with NCCL"""
    lib = load_library()
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    lib.smcb_obs_allreduce.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.smcb_obs_allreduce.restype = C.c_int
    rc = lib.smcb_obs_allreduce(arr, len(engines))
    if rc != 0:
        raise SmcbError(f"smcb error {rc}: {lib.smcb_last_error().decode()}")


def obs_layout_host(ngroups, nebins=64, e_lo=-8.0, e_hi=2.0):
    """the layout smcb_obs_layout_get reports, computed on the host (no engine needed): per group
    D[33^3] Mu[33^3] zprof[33] ehist[nebins] nsamples | sumE sumE2 sumP sumP2 sumAcc"""
    nvox, nz = 33 * 33 * 33, 33
    u64 = 2 * nvox + nz + nebins + 1
    return ObsLayout(ngroups, nvox, nz, nebins, e_lo, e_hi, u64, 5, u64 * ngroups, 5 * ngroups)


def unpack_obs(lay, cnt, mom):
    """split the packed observable block (smcb_obs_layout) into named arrays per group"""
    groups = []
    nv, nz, ne = lay.nvox, lay.nz, lay.nebins
    for g in range(lay.ngroups):
        c = cnt[g * lay.u64_per_group:(g + 1) * lay.u64_per_group]
        m = mom[g * lay.f64_per_group:(g + 1) * lay.f64_per_group]
        groups.append({
            "D": c[:nv].reshape(33, 33, 33), "Mu": c[nv:2 * nv].reshape(33, 33, 33),
            "zprof": c[2 * nv:2 * nv + nz], "ehist": c[2 * nv + nz:2 * nv + nz + ne],
            "nsamples": int(c[2 * nv + nz + ne]),
            "sumE": m[0], "sumE2": m[1], "sumP": m[2], "sumP2": m[3], "sumAcc": m[4],
        })
    return groups
