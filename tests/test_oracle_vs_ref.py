"""The CPU restatement (oracle/smc_oracle.c) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref, built by oracle/build_ref.sh).  Bit-for-bit: both are built with
-O2 -ffp-contract=off and the restatement keeps the reference's operation order.
Reference routines: SMC.c:278-351, 557-895, 912-927, 413-465; matematicose.c:183-193."""
import numpy as np
import pytest

from oracle_bindings import (GOLDEN_W_M3, RAND_MAX, Oracle, RefLib, RefNoWall, config_droplet, config_gas,
                             config_slab, make_sys, random_walls)

GEOM = {32: (33.0, 200.0), 108: (33.0, 200.0), 256: (33.0, 240.0), 500: (33.0, 240.0), 4096: (33.0, 240.0)}


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _configs(N, L, Lz, seed):
    rng = np.random.default_rng(seed)
    out = {"gas": config_gas(N, L, Lz, rng), "droplet": config_droplet(N, L, Lz, rng)}
    if N <= 256:
        out["slab"] = config_slab(N, L, Lz, rng)
    return out


@pytest.mark.parametrize("N,M", [(32, 3), (108, 3), (256, 3), (108, 4), (500, 3)])
def test_static_routines_bit_exact(orc, N, M):
    ref = RefLib(N, M)
    L, Lz = GEOM[N]
    s = make_sys(N, M, L, Lz)
    rng = np.random.default_rng(100 + N + M)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    cfgs = _configs(N, L, Lz, N * 7 + M)
    cfgs["lattice"] = ref.initializeBox(L, Lz)
    for name, R in cfgs.items():
        assert orc.energy(s, R) == ref.energy(R, L), name
        assert orc.pressure(s, R) == ref.pressure(R, L, Lz), name
        assert orc.walls_energy(s, R, W) == ref.wallsEnergy(R, W, L, Lz), name
        assert orc.walls_pressure(s, R, W) == ref.wallsPressure(R, W, L, Lz), name
        F0 = rng.standard_normal(3 * N)            # forces() accumulates into F
        np.testing.assert_array_equal(orc.forces(s, R, F0.copy()), ref.forces(R, L, F0.copy()), err_msg=name)
        for i in list(range(0, N, max(1, N // 16))) + [N - 1]:
            assert orc.energy_single(s, R, i) == ref.energySingle(R, L, i), (name, i)
            np.testing.assert_array_equal(orc.force_single(s, R, i), ref.forceSingle(R, L, i))
            p = R[3 * i:3 * i + 3]
            assert orc.walls_energy_single(s, p, W) == ref.wallsEnergySingle(p, W, L, Lz)
            f0 = rng.standard_normal(3)              # wallsForce adds
            np.testing.assert_array_equal(orc.walls_force(s, p, W, f0.copy()), ref.wallsForce(p, W, L, Lz, f0))


def test_static_N4096_droplet(orc):
    N = 4096
    ref = RefLib(N, 3)
    L, Lz = GEOM[N]
    s = make_sys(N, 3, L, Lz)
    rng = np.random.default_rng(5)
    R = config_droplet(N, L, Lz, rng, nz=8)
    W = GOLDEN_W_M3.copy()
    assert orc.energy(s, R) == ref.energy(R, L)
    assert orc.walls_energy(s, R, W) == ref.wallsEnergy(R, W, L, Lz)
    for i in (0, 17, 4095):
        np.testing.assert_array_equal(orc.force_single(s, R, i), ref.forceSingle(R, L, i))


def test_wall_clamp_outside_slab(orc):
    """particles at/after the walls take the dz=+-1e-4 clamp (SMC.c:738-739): ~1e40 energies"""
    N = 32
    ref = RefLib(N, 3)
    L, Lz = GEOM[N]
    s = make_sys(N, 3, L, Lz)
    W = GOLDEN_W_M3.copy()
    for z in (-Lz / 2, Lz / 2, -Lz / 2 - 3.0, Lz / 2 + 0.2, -Lz / 2 + 1e-9, Lz / 2 - 1e-9, 0.0):
        p = np.array([1.3, -4.0, z])
        assert orc.walls_energy_single(s, p, W) == ref.wallsEnergySingle(p, W, L, Lz)
        np.testing.assert_array_equal(orc.walls_force(s, p, W), ref.wallsForce(p, W, L, Lz))


def test_box_muller_and_initialize_box(orc):
    ref = RefLib(108, 3)
    rng = np.random.default_rng(3)
    rnd = rng.integers(0, RAND_MAX, size=324, endpoint=True, dtype=np.int64).astype(np.int32)
    rnd[:4] = [0, RAND_MAX, RAND_MAX, 0]
    ref.set_replay(rnd)
    a = ref.vecBoxMuller(1.4832396974191326, 324)
    assert ref.replay_pos() == 324 and ref.replay_underflow() == 0
    ref.set_replay(None)
    np.testing.assert_array_equal(a, orc.box_muller(1.4832396974191326, 324, rnd))
    for N in (32, 108, 256, 500):
        L, Lz = GEOM[N]
        X, sites = orc.initialize_box(L, Lz, N)
        assert sites == N
        np.testing.assert_array_equal(X, RefLib(N, 3).initializeBox(L, Lz))
    # SURVEY App. D anchors
    X, _ = orc.initialize_box(33.0, 240.0, 256)
    np.testing.assert_array_equal(X[:6], [2.0625, 2.0625, 2.0625, 6.1875, 6.1875, 2.0625])


@pytest.mark.parametrize("N,M,T,A,start,nsweeps", [(108, 3, 1.1, 1.1, "lattice", 60), (108, 3, 0.8, 0.004, "droplet", 40),
                                                   (256, 3, 1.1, 0.01, "droplet", 12), (32, 3, 1.1, 1.1, "slab", 100),
                                                   (108, 4, 1.1, 1.1, "gas", 30)])
def test_sweeps_bit_exact_with_replayed_stream(orc, N, M, T, A, start, nsweeps):
    """oneParticleMoves (SMC.c:278-351) fed a replayed rand() stream vs the restatement fed the same
    integers: positions, running energy and acceptance counts identical after every sweep."""
    ref = RefLib(N, M)
    L, Lz = GEOM[N]
    s = make_sys(N, M, L, Lz)
    rng = np.random.default_rng(N + 13 * M + nsweeps)
    W = GOLDEN_W_M3.copy() if M == 3 else random_walls(M, rng)
    if start == "lattice":
        R0 = ref.initializeBox(L, Lz)
    elif start == "droplet":
        R0 = config_droplet(N, L, Lz, rng, jitter=0.03)
    elif start == "slab":
        R0 = config_slab(N, L, Lz, rng)
    else:
        R0 = config_gas(N, L, Lz, rng)
    Ra, Rb = R0.copy(), R0.copy()
    Ea = Eb = orc.energy(s, R0) + orc.walls_energy(s, R0, W)
    per = 4 * N + 1
    stream = rng.integers(0, RAND_MAX, size=per * nsweeps, endpoint=True, dtype=np.int64).astype(np.int32)
    ref.set_replay(stream)
    tot_a = tot_b = 0
    for k in range(nsweeps):
        ja, Ea = ref.oneParticleMoves(Ra, W, L, Lz, A, T, Ea)
        jb, Eb = orc.sweep_from_ints(s, Rb, W, A, T, stream[k * per:(k + 1) * per], Eb)
        assert ja == jb, k
        assert Ea == Eb, k
        np.testing.assert_array_equal(Ra, Rb, err_msg=f"sweep {k}")
        tot_a += ja
    assert ref.replay_pos() == per * nsweeps and ref.replay_underflow() == 0
    ref.set_replay(None)
    assert 0 < tot_a < N * nsweeps or start == "lattice"
    # the running energy tracks the recomputed one (SURVEY §4 invariant)
    Erec = orc.energy(s, Rb) + orc.walls_energy(s, Rb, W)
    assert abs(Erec - Eb) <= 1e-9 * max(1.0, abs(Erec))


def test_local_density_matches(orc):
    N = 108
    ref = RefLib(N, 3)
    L, Lz = GEOM[N]
    s = make_sys(N, 3, L, Lz)
    rng = np.random.default_rng(8)
    Da, Ma = np.zeros(33 ** 3, dtype=np.uint64), np.zeros(33 ** 3, dtype=np.uint64)
    Db, Mb = Da.copy(), Ma.copy()
    Ba, Bb = np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
    for _ in range(5):
        R = config_gas(N, L, Lz, rng, zfrac=0.4999)
        ref.localDensityAndMobility(R, L, Lz, Da, Ba, Ma)
        orc.local_density(s, R, Db, Bb, Mb)
    np.testing.assert_array_equal(Da, Db)
    np.testing.assert_array_equal(Ma, Mb)
    np.testing.assert_array_equal(Ba, Bb)
    assert Da.sum() == 5 * N


def test_bulk_mode_matches_nowall_prototype(orc):
    """config 1: 3-D periodic bulk LJ, N=108, rho*=0.5 (L=6), cutoff L/2.  energy/forces/pressure of
    SMC_noMPI_noWall.c:573-591, 464-493, 664-684 are the pins (its single-particle routines are not)."""
    N, L = 108, (108 / 0.5) ** (1.0 / 3.0)
    ref = RefNoWall(N)
    s = make_sys(N, 3, L, L, rc2=L * L / 4, periodic_z=1, wall=0)
    rng = np.random.default_rng(7)
    R = ref.initializeBox(L) + (rng.random(3 * N) - 0.5) * 0.1
    assert orc.energy(s, R) == ref.energy(R, L)
    assert orc.pressure(s, R) == ref.pressure(R, L)
    np.testing.assert_array_equal(orc.forces(s, R), ref.forces(R, L))
    # internal consistency of the bulk extension: forceSingle == forces()[i] to rounding
    F = orc.forces(s, R)
    for i in (0, 1, 50, 107):
        np.testing.assert_allclose(orc.force_single(s, R, i), F[3 * i:3 * i + 3], rtol=1e-11, atol=1e-11)


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10"""
    np.testing.assert_array_equal(orc.philox([0, 0, 0, 0], [0, 0]),
                                  np.array([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], dtype=np.uint32))
    f = 0xffffffff
    np.testing.assert_array_equal(orc.philox([f, f, f, f], [f, f]),
                                  np.array([0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd], dtype=np.uint32))
    np.testing.assert_array_equal(orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
                                  np.array([0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1], dtype=np.uint32))
