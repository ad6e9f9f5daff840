"""CPU parity of the drop-in's host-side post-processing against the compiled reference (oracle/_ref):
fft_acf (SMC.c:1055-1090; FFTW3 backed by the oracle's definition-level DFT shim) and clusterAnalysis
(SMC.c:971-1045), quirks included.  No GPU is touched: these routines are host code in both builds."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_bindings import RefLib, config_droplet

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "dropin", "_build")


class DoubleArray(C.Structure):
    _fields_ = [("length", C.c_size_t), ("data", C.POINTER(C.c_double))]


def _libs(N=108):
    path = os.path.join(BUILD, f"libdropin_N{N}_M3.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run `make -C tests/dropin` (or __graft_entry__.build())")
    ref, drop = RefLib(N, 3), RefLib(N, 3, path=path)
    for lib in (ref.lib, drop.lib):
        lib.fft_acf.restype = DoubleArray
        lib.fft_acf.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64), C.c_size_t, C.c_int]
        lib.clusterAnalysis.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64), C.c_int, C.c_double,
                                        np.ctypeslib.ndpointer(dtype=np.int32)]
        lib.simple_acf.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64), C.c_size_t, C.c_int,
                                   np.ctypeslib.ndpointer(dtype=np.float64)]
    return ref, drop


def _acf(lib, H, kmax):
    out = lib.fft_acf(H, H.size, kmax)
    v = np.array([out.data[k] for k in range(out.length)])
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(out.data, C.c_void_p))
    return v


def _series(n, seed):
    """an AR(1)-like energy series with a slow drift, like sMC's E[]"""
    rng = np.random.default_rng(seed)
    x = np.empty(n)
    x[0] = -10.0
    eps = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = -10.0 + 0.97 * (x[i - 1] + 10.0) + 0.3 * eps[i]
    return x + 1e-3 * np.arange(n)


@pytest.mark.parametrize("n,kmax", [(2001, 300), (2000, 300), (1537, 100), (64, 40), (401, 2500000)])
def test_fft_acf_matches_the_reference(n, kmax):
    """same values to 1e-10 for odd / even lengths, a power of two, and the k_max reduction of short series
    (sMC always asks for KMAX = 2 500 000 lags, SMC.h:61)"""
    ref, drop = _libs()
    H = _series(n, seed=n)
    a, b = _acf(ref.lib, H, kmax), _acf(drop.lib, H, kmax)
    assert a.shape == b.shape and a.size == (kmax if n >= 2 * kmax + 1 else int(np.rint(n // 2)) - 2)
    assert abs(b[0] - 1.0) < 1e-14
    lfft = n // 2 + n % 2
    m = min(a.size, lfft)                      # the reference reads past its transform beyond lfft lags
    np.testing.assert_allclose(b[:m], a[:m], rtol=0, atol=1e-10)
    # and it is what the formula says: Re sum_j |F_j|^2 e^{2 pi i jk/lfft} / sum_j |F_j|^2 over the first lfft bins
    F = np.fft.fft(H - H.mean())[:lfft]
    want = np.real(np.fft.ifft(np.abs(F) ** 2) * lfft)
    np.testing.assert_allclose(b[:m], (want / want[0])[:m], rtol=0, atol=1e-10)


def test_fft_acf_long_series_is_not_truncated():
    """the first-generation drop-in capped k_max so its direct sums stayed affordable; the O(n log n) transform
    does a quarter-million-point series whole (the reference's production runs hold 16e6 sweeps)"""
    ref, drop = _libs()
    n, kmax = 262145, 100000
    H = _series(n, seed=5)
    b = _acf(drop.lib, H, kmax)
    assert b.size == kmax
    lfft = n // 2 + 1
    F = np.fft.fft(H - H.mean())[:lfft]
    want = np.real(np.fft.ifft(np.abs(F) ** 2) * lfft)
    np.testing.assert_allclose(b, (want / want[0])[:kmax], rtol=0, atol=1e-9)


def test_simple_acf_matches_the_reference():
    ref, drop = _libs()
    H = _series(1500, seed=3)
    a, b = np.zeros(200), np.zeros(200)
    ref.lib.simple_acf(H, H.size, 200, a)
    drop.lib.simple_acf(H, H.size, 200, b)
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-13)


@pytest.mark.parametrize("kind,seed", [("droplet", 1), ("droplet", 2), ("gas", 3)])
def test_cluster_analysis_is_the_reference_array(kind, seed):
    """LCA[3*npairs] exactly as the reference fills it (its overlapping slot arithmetic included) on condensed
    droplets - where there are bonds, shared neighbours and chains to count - and on a gas"""
    N, L, Lz = 108, 33.0, 200.0
    ref, drop = _libs(N)
    rng = np.random.default_rng(seed)
    if kind == "droplet":
        R = config_droplet(N, L, Lz, rng, jitter=0.08, nz=4)
        R = R + 0.1 * rng.standard_normal(R.shape)
    else:
        R = (rng.random(3 * N) - 0.5) * np.tile([L, L, 20.0], N)
    n3 = 3 * (N * N - N) // 2
    a, b = np.full(n3, -7, dtype=np.int32), np.full(n3, -7, dtype=np.int32)
    ref.lib.clusterAnalysis(np.ascontiguousarray(R), N, L, a)
    drop.lib.clusterAnalysis(np.ascontiguousarray(R), N, L, b)
    np.testing.assert_array_equal(b, a)
    if kind == "droplet":
        assert a[0::3].sum() > N and a[1::3].max() >= 2 and a[2::3].max() >= 1     # the case is not trivial
