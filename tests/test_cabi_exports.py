"""The C-ABI library loads and exports every symbol include/smcb200.h declares; without a CUDA device
the product refuses to run (no CPU fallback) instead of computing anything."""
import ctypes
import importlib
import os
import subprocess

smcb = importlib.import_module("montecarlo-surfacer_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = smcb.header_symbols()
    assert len(hdr) >= 30 and "smcb_sweep" in hdr and "smcb_step_allparticle" in hdr
    assert sorted(smcb.exported_symbols()) == hdr


def test_library_is_sm100a_only_and_does_not_link_the_oracle():
    so = smcb.lib_path()
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs
    needed = subprocess.run(["objdump", "-p", so], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "libref" not in needed
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert " orc_" not in syms


def test_no_device_means_error_not_fallback():
    lib = smcb.load_library()
    h = ctypes.c_void_p()
    import torch
    rc = lib.smcb_create(ctypes.byref(h), 0, 4, 32, 3)
    if torch.cuda.is_available():
        assert rc == 0
        lib.smcb_destroy(h)
    else:
        assert rc == -3 and not h.value                      # SMCB_ERR_NODEVICE
        assert b"no CPU path" in lib.smcb_last_error()
    assert lib.smcb_create(ctypes.byref(h), 0, 0, 32, 3) == -1   # SMCB_ERR_ARG before any device work


def test_dropin_exports_the_reference_api():
    """the drop-in library (the reference's SMC.h API on libsmcb200, built by tests/dropin/Makefile) defines every
    function the reference's SMC.h:92-121 and matematicose.h:6-28 declare; no GPU needed to check that"""
    so = os.path.join(ROOT, "tests", "dropin", "_build", "libdropin_N108_M3.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "dropin")], check=True, capture_output=True)
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in syms.splitlines() if ln.strip()}
    smc_h = ["sMC", "vecBoxMuller", "shiftSystem", "shiftSystem2D", "shiftSystem3D", "createZRange", "initializeWalls",
             "initializeBox", "oneParticleMoves", "energySingle", "forceSingle", "forces", "energy", "pressure", "wallsEnergy",
             "wallsEnergySingle", "wallsForce", "wallsPressure", "localDensityAndMobility", "localDensityAndMobility_nonuniz",
             "clusterAnalysis", "boundsCheck", "simple_acf", "fft_acf", "variance_corr"]
    mat_h = ["isPicoEqual", "pointwise", "double_max_index", "double_min_index", "sum", "intsum", "mean", "intmean", "variance",
             "zeros", "elforel", "isApproxEqual", "zerosecant", "secant", "findzero_last", "fast_bessel", "der3", "der5", "der5_c",
             "simpson_integral", "grad_descent_1D", "stochastic_grad_descent_1D"]
    missing = [f for f in smc_h + mat_h if f not in defined]
    assert not missing, missing
    needed = subprocess.run(["objdump", "-p", so], capture_output=True, text=True).stdout
    assert "libsmcb200.so" in needed and "fftw" not in needed.lower()
