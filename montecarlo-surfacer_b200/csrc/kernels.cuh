// kernels.cuh — the sm_100a kernels of libsmcb200 (templates; instantiated in
// kernels_fast.cu with FMA contraction and in kernels_strict.cu with --fmad=false).
//
//   k_evaluate      static per-particle / per-chain energies, forces, virials
//                   (reference rows a2-a10: SMC.c:557-895)
//   k_sweep         the reference's live move, oneParticleMoves (SMC.c:278-351), in
//                   the reference's arithmetic (STRICT, bit-exact parity): N sequential
//                   single-particle Smart-MC trials, one WARP per chain, positions
//                   register-resident.  The FAST sweep is k_sweep_cached (sweep_cached.cuh),
//                   the FAST all-particle step and evaluation are in allparticle_fast.cuh
//   k_allparticle   the all-particle Smart-MC step (north-star kernel B), one CTA
//                   per chain, proposal positions tiled in shared memory
//   k_gather        localDensityAndMobility (SMC.c:912-927) + histograms/moments
//
// Device layout (HBM): positions are SoA per chain, pos[chain][3][Npad] doubles
// (Npad = N rounded up to 32) so a warp's load of 32 consecutive particles is one
// 256-byte coalesced transaction per component.
#pragma once
#include "smcb_common.cuh"
#include "philox.cuh"

namespace smcb {

struct DevChains {
    int C, N, Npad, M;
    const smcb_chain_params *params;
    int nparams;
    const double *W;          // [nwalls][2*M*M]
    double *pos;              // [C][3][Npad]
    double *E;                // [C] running potential energy
    long long *nacc, *ntri;   // [C]
    double step_scale;        // A multiplier (thermalisation uses 2, SMC.c:110)
    const float *extent;      // nullable [C][2]: max |x|/L,|y|/L and max |z|/L of the chain at upload (FP32 screen bound)
    unsigned long long *pair_counts;   // [0] ordered pair-interactions as the reference executes them (nominal), [1] of them inside
                                       // the cutoff, [2] pair distance tests the kernel actually EXECUTED (cached / screened / half-shell
                                       // kernels execute fewer than [0])
};

constexpr int kTot = 5;      // chain totals: U_lj, U_wall, vir_lj, vir_wall_ref (as the reference writes it), vir_wall (as it meant it)

struct EvalOut {              // all nullable, SoA
    double *e_lj, *f_lj, *e_wall, *f_wall;   // [C][Npad], [C][3][Npad]
    double *totals;                          // [C][kTot]
};

struct RngArgs {
    uint32_t k0, k1, chain0;
    unsigned long long step0;
};

struct SweepArgs {
    int nsweeps;
    RngArgs rng;
    const double *displ;       // FED: [s][C][3N]
    const long long *offset;   // FED: [s][C]
    const double *u;           // FED: [s][C][N]
    unsigned char *accepted;   // FED, nullable: [s][C][N]
    double *cache_out;         // test hook of the cached kernel, nullable: [C][5][Npad]
    double *trace_E;           // nullable [s][C]: running energy after every sweep (sMC's E[n+1], SMC.c:116,194)
    int *trace_acc;            // nullable [s][C]: accepted trials of every sweep (sMC's jj[n])
    int refresh_E;             // 1: d.E is stale - the FAST warp kernels take the chain energy from their cache rebuild (SMC.c:48)
    int dense_hint;            // host side only: 1 = the previous launch found > 2 % of the pairs inside the cutoff (condensed phase)
};

struct StepArgs {
    int nsteps;
    int refresh;               // 1: F and E are stale, recompute them from pos first
    RngArgs rng;
    double *F, *Fn, *dl;       // [C][3][Npad] current force, proposal force, displacement
    const double *xi;          // FED: [s][C][3N], already scaled by sqrt(2A)
    const double *u;           // FED: [s][C]
    double *lnap;              // nullable [s][C]
    unsigned char *accepted;   // nullable [s][C]
};

__device__ __forceinline__ const smcb_chain_params &chain_params(const DevChains &d, int chain)
{
    return d.params[d.nparams == 1 ? 0 : chain];
}

// ============================================================ k_evaluate ====
// One CTA per chain; thread i owns particle i (strided) and walks l = 0..N-1 in
// ascending order over the shared-memory copy, so in STRICT mode every
// per-particle sum is formed in exactly the reference's order.
template <bool STRICT>
__device__ __forceinline__ void particle_vs_all(const Box &b, const double *__restrict__ W, int N, int i,
                                                const double *sx, const double *sy, const double *sz,
                                                double &e_lj, double &e_wall, double &fx, double &fy, double &fz,
                                                double &wx, double &wy, double &wz, double *vir, unsigned long long &cnt)
{
    const double xi = sx[i], yi = sy[i], zi = sz[i];
    double e = 0.0, v = 0.0;
    fx = fy = fz = 0.0;
#pragma unroll 4
    for (int l = 0; l < N; l++) {
        double dx, dy, dz;
        const double r2 = pair_sep<STRICT>(b, xi, yi, zi, sx[l], sy[l], sz[l], dx, dy, dz);
        if (r2 < b.rc2 && l != i) {
            double et, g;
            lj_terms<STRICT, true>(r2, 1.0, 1.0, et, g);
            e += et;
            fx += g * dx;
            fy += g * dy;
            fz += g * dz;
            if (vir) v += virial_term<STRICT>(r2);
            cnt++;
        }
    }
    e_lj = e * 4;
    if (vir) *vir = v;
    wx = wy = wz = 0.0;
    e_wall = 0.0;
    if (b.wall) e_wall = wall_particle<STRICT>(b, W, xi, yi, zi, wx, wy, wz) * 4;
}

template <bool STRICT>
__global__ void k_evaluate(DevChains d, EvalOut o)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad;
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + Npad, *sz = sm + 2 * Npad, *scratch = sm + 3 * Npad;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    const double *P = d.pos + (size_t)chain * 3 * Npad;
    for (int j = threadIdx.x; j < Npad; j += blockDim.x) {
        sx[j] = P[j]; sy[j] = P[Npad + j]; sz[j] = P[2 * Npad + j];
    }
    __syncthreads();
    double tot[kTot] = {0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned long long cnt = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double e_lj, e_wall, fx, fy, fz, wx, wy, wz, vir;
        particle_vs_all<STRICT>(b, W, N, i, sx, sy, sz, e_lj, e_wall, fx, fy, fz, wx, wy, wz, &vir, cnt);
        const size_t q = (size_t)chain * Npad + i, q3 = (size_t)chain * 3 * Npad + i;
        if (o.e_lj) o.e_lj[q] = e_lj;
        if (o.e_wall) o.e_wall[q] = e_wall;
        if (o.f_lj) { o.f_lj[q3] = fx; o.f_lj[q3 + Npad] = fy; o.f_lj[q3 + 2 * Npad] = fz; }
        if (o.f_wall) { o.f_wall[q3] = wx; o.f_wall[q3 + Npad] = wy; o.f_wall[q3 + 2 * Npad] = wz; }
        tot[0] += 0.5 * e_lj;
        tot[1] += e_wall;
        tot[2] += 0.5 * vir;
        if (b.wall) {
            tot[3] += wall_virial_ref<STRICT>(b, W, sx[i], sy[i], sz[i]);
            tot[4] += wall_virial_intended<STRICT>(b, W, sx[i], sy[i], sz[i]);
        }
    }
    block_sum<kTot>(tot, scratch);
    if (threadIdx.x == 0 && o.totals) {
        double *t = o.totals + (size_t)chain * kTot;
        for (int k = 0; k < kTot; k++) t[k] = tot[k];
    }
}

// ================================================================ k_sweep ===
// One warp per chain.  Lane `lane` keeps particles j = lane + 32k (k < K) in
// registers for the whole launch; a shared-memory mirror serves the broadcast
// read of the trial particle.  A trial is two passes over the K register slots
// (old position, proposed position), each followed by a 4-value warp reduction.
//
// STRICT: the warp forms every sum in the reference's order.  Lanes compute the
// pair terms in parallel; a ballot finds the in-cutoff lanes and their terms are
// added one by one in ascending particle index (k-major, lane-minor = l
// ascending), all lanes keeping identical accumulators.  The surface terms
// follow in the reference's order (flat wall, then sites m ascending).
// ---- shared FAST-path helpers (used by sweep_cached.cuh and allparticle_fast.cuh) ----------------
// 1/x for a positive normal x to ~1 ulp without the slow-path call of the IEEE
// division: MUFU.RCP64H seed (rcp.approx.ftz.f64, 2^-23) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// Sum four per-lane values over the warp with a transposed butterfly: the first two
// exchange steps halve the number of live values instead of reducing each of the four
// separately (20 SHFL + 6 DADD instead of 40 + 20).  Every lane returns all four totals.
__device__ __forceinline__ void warp_sum4(int lane, double &a, double &b, double &c, double &d)
{
    const bool hi = (lane & 16) != 0;
    double k0 = hi ? c : a, k1 = hi ? d : b;          // kept pair
    const double s0 = hi ? a : c, s1 = hi ? b : d;    // sent pair
    k0 += __shfl_xor_sync(FULL, s0, 16);
    k1 += __shfl_xor_sync(FULL, s1, 16);
    const bool mid = (lane & 8) != 0;
    double v = mid ? k1 : k0;
    const double w = mid ? k0 : k1;
    v += __shfl_xor_sync(FULL, w, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    a = __shfl_sync(FULL, v, 0);
    b = __shfl_sync(FULL, v, 8);
    c = __shfl_sync(FULL, v, 16);
    d = __shfl_sync(FULL, v, 24);
}

template <int K, bool STRICT>
__device__ __forceinline__ void sweep_pass(const Box &b, const double *__restrict__ W, int N, int lane,
                                           int owner, int slot, double px, double py, double pz,
                                           const double (&x)[K], const double (&y)[K], const double (&z)[K],
                                           double &U, double &Fx, double &Fy, double &Fz, unsigned long long &cnt)
{
    const int MM = b.M * b.M;
    const double dw = b.L / b.M;
    if (STRICT) {
        double V = 0.0;
        Fx = Fy = Fz = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            double dx, dy, dz;
            const double r2 = pair_sep<true>(b, px, py, pz, x[k], y[k], z[k], dx, dy, dz);
            const bool ok = (r2 < b.rc2) && (lane + 32 * k < N) && !(k == slot && lane == owner);
            unsigned mask = __ballot_sync(FULL, ok);
            if (mask) {
                double et = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
                if (ok) {
                    double g;
                    lj_terms<true, true>(r2, 1.0, 1.0, et, g);
                    gx = g * dx; gy = g * dy; gz = g * dz;
                    cnt++;
                }
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    V += __shfl_sync(FULL, et, src);
                    Fx += __shfl_sync(FULL, gx, src);
                    Fy += __shfl_sync(FULL, gy, src);
                    Fz += __shfl_sync(FULL, gz, src);
                }
            }
        }
        double Vw = 0.0;
        if (b.wall) {
            const double dzw = wall_dz<true>(b, pz);
            double e0, g0;
            zwall_terms<true>(b, dzw, e0, g0);
            Vw += e0;
            Fz += g0 * dzw;
            for (int m0 = 0; m0 < MM; m0 += 32) {
                const int m = m0 + lane;
                const int i = m / b.M, j = m - i * b.M;
                const double dx = min_image<true>(px - i * dw, b.L, b.invL);
                const double dy = min_image<true>(py - j * dw, b.L, b.invL);
                const double r2 = dx * dx + dy * dy + dzw * dzw;
                const bool ok = (m < MM) && (r2 < b.rc2);
                unsigned mask = __ballot_sync(FULL, ok);
                if (mask) {
                    double et = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
                    if (ok) {
                        double g;
                        lj_terms<true, false>(r2, W[2 * m], W[2 * m + 1], et, g);
                        gx = g * dx; gy = g * dy; gz = g * dzw;
                    }
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        Vw += __shfl_sync(FULL, et, src);
                        Fx += __shfl_sync(FULL, gx, src);
                        Fy += __shfl_sync(FULL, gy, src);
                        Fz += __shfl_sync(FULL, gz, src);
                    }
                }
            }
        }
        U = V * 4 + Vw * 4;          // energySingle(...) + wallsEnergySingle(...)  (SMC.c:300)
    } else {
        double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            double dx, dy, dz;
            const double r2 = pair_sep<false>(b, px, py, pz, x[k], y[k], z[k], dx, dy, dz);
            const bool ok = (r2 < b.rc2) && (lane + 32 * k < N) && !(k == slot && lane == owner);
            if (ok) {
                double et, g;
                lj_terms<false, true>(r2, 1.0, 1.0, et, g);
                e += et;
                fx = fma(g, dx, fx);
                fy = fma(g, dy, fy);
                fz = fma(g, dz, fz);
                cnt++;
            }
        }
        if (b.wall) {
            const double dzw = wall_dz<false>(b, pz);
            for (int m = lane; m < MM; m += 32) {
                const int i = m / b.M, j = m - i * b.M;
                const double dx = min_image<false>(px - i * dw, b.L, b.invL);
                const double dy = min_image<false>(py - j * dw, b.L, b.invL);
                const double r2 = fma(dzw, dzw, fma(dy, dy, dx * dx));
                if (r2 < b.rc2) {
                    double et, g;
                    lj_terms<false, false>(r2, W[2 * m], W[2 * m + 1], et, g);
                    e += et;
                    fx = fma(g, dx, fx);
                    fy = fma(g, dy, fy);
                    fz = fma(g, dzw, fz);
                }
            }
            if (lane == 31) {        // the flat wall rides on the lane the site loop uses last
                double e0, g0;
                zwall_terms<false>(b, dzw, e0, g0);
                e += e0;
                fz = fma(g0, dzw, fz);
            }
        }
        U = 4.0 * warp_sum(e);
        Fx = warp_sum(fx);
        Fy = warp_sum(fy);
        Fz = warp_sum(fz);
    }
}

template <int K, bool STRICT, bool FED>
__global__ void __launch_bounds__(32, (K <= 8 ? 16 : 8)) k_sweep(DevChains d, SweepArgs a)
{
    const int lane = threadIdx.x, chain = blockIdx.x;
    const int N = d.N, Npad = d.Npad;
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + Npad, *sz = sm + 2 * Npad;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;

    // registers: the lane's particles, unscaled (STRICT) or in units of L (FAST)
    double x[K], y[K], z[K];
    unsigned validmask = 0;
    const double rscale = STRICT ? 1.0 : b.invL;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        const bool in = j < N;
        const double X = in ? P[j] : 0.0, Y = in ? P[Npad + j] : 0.0, Z = in ? P[2 * Npad + j] : 0.0;
        if (j < Npad) { sx[j] = X; sy[j] = Y; sz[j] = Z; }
        x[k] = X * rscale; y[k] = Y * rscale; z[k] = Z * rscale;
        if (in) validmask |= 1u << k;
    }
    __syncwarp();
    static_assert(STRICT, "k_sweep is the parity kernel; the FAST sweep is k_sweep_cached (sweep_cached.cuh)");
    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    double E = d.E[chain];
    long long nacc = 0;
    unsigned long long cnt = 0;
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};

    for (int s = 0; s < a.nsweeps; s++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)s;
        const size_t sc = (size_t)s * d.C + chain;
        const long long nacc0 = nacc;
        long long offset;                              // int offset = rand();  SMC.c:290
        if (FED) {
            offset = a.offset[sc];
        } else {
            uint32_t o; double unused;
            rng_step_scalars(id, step, o, unused);
            offset = o;
        }
        const int off = (int)(offset % N);
        for (int nn0 = 0; nn0 < N; nn0 += 32) {
            // each lane prepares the random inputs of one of the next 32 trials
            const int nnl = nn0 + lane;
            int nl = nnl + off;                        // n = (nn+offset)%N  SMC.c:294
            if (nl >= N) nl -= N;
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, ul = 2.0;
            if (nnl < N) {
                if (FED) {
                    const double *dsp = a.displ + sc * 3 * N;
                    g0 = dsp[3 * nl]; g1 = dsp[3 * nl + 1]; g2 = dsp[3 * nl + 2];
                    ul = a.u[sc * N + nnl];
                } else {
                    rng_particle_gauss(id, step, (uint32_t)nl, g0, g1, g2);
                    g0 *= sigma; g1 *= sigma; g2 *= sigma;
                    ul = rng_particle_uniform(id, step, (uint32_t)nl);
                }
            }
            const int tmax = min(32, N - nn0);
            for (int t = 0; t < tmax; t++) {
                const int n = __shfl_sync(FULL, nl, t);
                const double gx = __shfl_sync(FULL, g0, t);
                const double gy = __shfl_sync(FULL, g1, t);
                const double gz = __shfl_sync(FULL, g2, t);
                const double uu = __shfl_sync(FULL, ul, t);
                const int owner = n & 31, slot = n >> 5;
                const double px = sx[n], py = sy[n], pz = sz[n];

                double Um, Fmx, Fmy, Fmz;              // SMC.c:300-304
                sweep_pass<K, STRICT>(b, W, N, lane, owner, slot, px, py, pz, x, y, z, Um, Fmx, Fmy, Fmz, cnt);

                double dX, dY, dZ;                      // SMC.c:307-309
                if (STRICT) {
                    dX = Fmx * b.A / b.T + gx;
                    dY = Fmy * b.A / b.T + gy;
                    dZ = Fmz * b.A / b.T + gz;
                } else {
                    dX = fma(Fmx, AoT, gx);
                    dY = fma(Fmy, AoT, gy);
                    dZ = fma(Fmz, AoT, gz);
                }
                double qx = px + dX, qy = py + dY, qz = pz + dZ;   // SMC.c:311-316
                qx = min_image<STRICT>(qx, b.L, b.invL);
                qy = min_image<STRICT>(qy, b.L, b.invL);
                if (b.pz) qz = min_image<STRICT>(qz, b.Lz, b.invLz);

                double Un, Fnx, Fny, Fnz;              // SMC.c:319-321
                sweep_pass<K, STRICT>(b, W, N, lane, owner, slot, qx, qy, qz, x, y, z, Un, Fnx, Fny, Fnz, cnt);

                double ap;                              // SMC.c:326-329
                if (STRICT) {
                    const double hx = Fnx - Fmx, hy = Fny - Fmy, hz = Fnz - Fmz;
                    const double dWk = (hx * hx + hy * hy + hz * hz + 2.0 * (hx * Fmx + hy * Fmy + hz * Fmz)) * b.A / (4.0 * b.T);
                    ap = exp(-(Un - Um + (dX * (Fnx + Fmx) + dY * (Fny + Fmy) + dZ * (Fnz + Fmz)) / 2.0 + dWk) / b.T);
                } else {
                    const double f2 = fma(Fnx, Fnx, fma(Fny, Fny, Fnz * Fnz)) - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                    const double dr = fma(dX, Fnx + Fmx, fma(dY, Fny + Fmy, dZ * (Fnz + Fmz)));
                    ap = exp(-((Un - Um) + 0.5 * dr + f2 * (0.25 * AoT)) / b.T);
                }
                const bool acc = uu < ap;               // SMC.c:335
                if (acc) {
                    if (lane == owner) {
                        sx[n] = qx; sy[n] = qy; sz[n] = qz;
#pragma unroll
                        for (int k = 0; k < K; k++)
                            if (k == slot) { x[k] = qx * rscale; y[k] = qy * rscale; z[k] = qz * rscale; }
                    }
                    E += Un - Um;                       // SMC.c:341
                    nacc++;
                }
                if (FED && a.accepted != nullptr && lane == 0) a.accepted[sc * N + nn0 + t] = acc ? 1 : 0;
                __syncwarp();
            }
        }
        if (a.trace_E != nullptr && lane == 0) { a.trace_E[sc] = E; a.trace_acc[sc] = (int)(nacc - nacc0); }
    }

    for (int j = lane; j < N; j += 32) {               // the shared-memory mirror holds the exact positions
        P[j] = sx[j]; P[Npad + j] = sy[j]; P[2 * Npad + j] = sz[j];
    }
    unsigned long long tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
    if (lane == 0) {
        d.E[chain] = E;
        d.nacc[chain] += nacc;
        d.ntri[chain] += (long long)a.nsweeps * N;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
            atomicAdd(d.pair_counts + 1, tot);
            atomicAdd(d.pair_counts + 2, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));    // two full passes per trial
        }
    }
}

}  // namespace smcb
#include "sweep_cached.cuh"
#include "sweep_spec.cuh"
#include "allparticle_fast.cuh"
#include "sweep_block.cuh"
#include "sweep_block_spec.cuh"
#include "sweep_block_strict.cuh"
#ifdef SMCB_MISC_KERNELS      // non-template kernels: one definition, in kernels_fast.cu
#include "fp32_mode.cuh"
#endif
namespace smcb {

// ========================================================== k_allparticle ===
// One CTA per chain, persistent over nsteps.  Shared memory holds the current
// and the proposed configuration (2 x 3 x Npad doubles); forces live in three
// L2-resident global arrays touched once per particle per step.
template <bool STRICT, bool FED>
__global__ void k_allparticle(DevChains d, StepArgs a)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad, tid = threadIdx.x;
    extern __shared__ double sm[];
    double *cur = sm, *nxt = sm + 3 * Npad, *scratch = sm + 6 * Npad;
    __shared__ int s_accept;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    double *Fc = a.F + (size_t)chain * 3 * Npad;
    double *Fn = a.Fn + (size_t)chain * 3 * Npad;
    double *DL = a.dl + (size_t)chain * 3 * Npad;
    for (int j = tid; j < 3 * Npad; j += blockDim.x) cur[j] = P[j];
    __syncthreads();

    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    unsigned long long cnt = 0;
    double U = d.E[chain];
    long long nacc = 0;

    if (a.refresh) {           // bring F and U in line with the positions
        double t[1] = {0.0};
        for (int i = tid; i < N; i += blockDim.x) {
            double e_lj, e_wall, fx, fy, fz, wx, wy, wz;
            particle_vs_all<STRICT>(b, W, N, i, cur, cur + Npad, cur + 2 * Npad, e_lj, e_wall, fx, fy, fz, wx, wy, wz, nullptr, cnt);
            Fc[i] = fx + wx; Fc[Npad + i] = fy + wy; Fc[2 * Npad + i] = fz + wz;
            t[0] += 0.5 * e_lj + e_wall;
        }
        block_sum<1>(t, scratch);
        U = t[0];
        cnt = 0;
    }

    for (int s = 0; s < a.nsteps; s++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)s;
        const size_t sc = (size_t)s * d.C + chain;
        // ---- proposal: d_i = F_i A/T + xi_i ; r' = wrap(r + d) ---------------
        for (int i = tid; i < N; i += blockDim.x) {
            double g0, g1, g2;
            if (FED) {
                const double *xi = a.xi + sc * 3 * N;
                g0 = xi[3 * i]; g1 = xi[3 * i + 1]; g2 = xi[3 * i + 2];
            } else {
                rng_particle_gauss(id, step, (uint32_t)i, g0, g1, g2);
                g0 *= sigma; g1 *= sigma; g2 *= sigma;
            }
            double dX, dY, dZ;
            if (STRICT) {
                dX = Fc[i] * b.A / b.T + g0;
                dY = Fc[Npad + i] * b.A / b.T + g1;
                dZ = Fc[2 * Npad + i] * b.A / b.T + g2;
            } else {
                dX = fma(Fc[i], AoT, g0);
                dY = fma(Fc[Npad + i], AoT, g1);
                dZ = fma(Fc[2 * Npad + i], AoT, g2);
            }
            DL[i] = dX; DL[Npad + i] = dY; DL[2 * Npad + i] = dZ;
            double qx = cur[i] + dX, qy = cur[Npad + i] + dY, qz = cur[2 * Npad + i] + dZ;
            qx = min_image<STRICT>(qx, b.L, b.invL);
            qy = min_image<STRICT>(qy, b.L, b.invL);
            if (b.pz) qz = min_image<STRICT>(qz, b.Lz, b.invLz);
            nxt[i] = qx; nxt[Npad + i] = qy; nxt[2 * Npad + i] = qz;
        }
        __syncthreads();
        // ---- forces and energy at the proposal, MH sums ----------------------
        double t[3] = {0.0, 0.0, 0.0};     // U', sum d.(F'+F), sum |F'|^2-|F|^2
        for (int i = tid; i < N; i += blockDim.x) {
            double e_lj, e_wall, fx, fy, fz, wx, wy, wz;
            particle_vs_all<STRICT>(b, W, N, i, nxt, nxt + Npad, nxt + 2 * Npad, e_lj, e_wall, fx, fy, fz, wx, wy, wz, nullptr, cnt);
            fx += wx; fy += wy; fz += wz;
            Fn[i] = fx; Fn[Npad + i] = fy; Fn[2 * Npad + i] = fz;
            const double ox = Fc[i], oy = Fc[Npad + i], oz = Fc[2 * Npad + i];
            t[0] += 0.5 * e_lj + e_wall;
            t[1] += DL[i] * (fx + ox) + DL[Npad + i] * (fy + oy) + DL[2 * Npad + i] * (fz + oz);
            t[2] += (fx * fx - ox * ox) + (fy * fy - oy * oy) + (fz * fz - oz * oz);
        }
        block_sum<3>(t, scratch);
        const double lnap = -((t[0] - U) + t[1] / 2.0 + t[2] * b.A / (4.0 * b.T)) / b.T;
        if (tid == 0) {
            double uu;
            if (FED) uu = a.u[sc];
            else { uint32_t o; rng_step_scalars(id, step, o, uu); }
            const int acc = uu < exp(lnap);
            s_accept = acc;
            if (a.lnap) a.lnap[sc] = lnap;
            if (a.accepted) a.accepted[sc] = (unsigned char)acc;
        }
        __syncthreads();
        if (s_accept) {
            double *tp = cur; cur = nxt; nxt = tp;
            tp = Fc; Fc = Fn; Fn = tp;
            U = t[0];
            nacc++;
        }
        __syncthreads();
    }

    for (int j = tid; j < 3 * Npad; j += blockDim.x) P[j] = cur[j];
    double *Fcanon = a.F + (size_t)chain * 3 * Npad;
    if (Fc != Fcanon)
        for (int j = tid; j < 3 * Npad; j += blockDim.x) Fcanon[j] = Fc[j];
    double c[1] = {(double)cnt};
    block_sum<1>(c, scratch);
    if (tid == 0) {
        d.E[chain] = U;
        d.nacc[chain] += nacc;
        d.ntri[chain] += a.nsteps;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsteps * (unsigned long long)N * (N - 1));
            atomicAdd(d.pair_counts + 1, (unsigned long long)c[0]);
            atomicAdd(d.pair_counts + 2, (unsigned long long)(a.nsteps + (a.refresh ? 1 : 0)) * (unsigned long long)N * (N - 1));
        }
    }
}

// =============================================================== k_gather ===
struct GatherArgs {
    const double *totals;          // [C][kTot] from the FAST evaluation
    int wall_virial_intended;      // 0: P uses the reference's wallsPressure arithmetic (default), 1: the intended wall virial
    int *rbin;                     // [C][N]
    unsigned long long *counters;  // [G][u64_per_group]
    double *moments;               // [G][f64_per_group]
    double *chain_mom;             // [C][5] this gather's contribution of every chain (summed per group in fixed order)
    int ngroups;
    size_t u64_per_group, f64_per_group;
    int nebins;
    double e_lo, e_hi;
};

// The kernels below are not templates: they are compiled once, in kernels_fast.cu.
#ifdef SMCB_MISC_KERNELS
// localDensityAndMobility (SMC.c:912-927) for every chain into its group's
// voxel block, plus the z profile, the energy histogram and the moments sMC
// accumulates at a gather (SMC.c:137-141).  Counters are exact (integer
// atomics, order-independent); the floating-point moments are written per chain and
// summed per group in a fixed order by k_gather_moments, so the whole observable
// block is bit-reproducible.
__global__ void k_gather(DevChains d, GatherArgs g)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad;
    const smcb_chain_params &cp = chain_params(d, chain);
    const double *P = d.pos + (size_t)chain * 3 * Npad;
    unsigned long long *cnt = g.counters + (size_t)cp.group * g.u64_per_group;
    const int nvox = SMCB_NCX * SMCB_NCX * SMCB_NCZ;
    unsigned long long *D = cnt, *Mu = cnt + nvox, *zprof = cnt + 2 * nvox, *ehist = zprof + SMCB_NCZ,
                       *nsamp = ehist + g.nebins;
    __shared__ unsigned zloc[SMCB_NCZ];            // the chain's z profile first: 33 hot addresses would serialise in L2
    for (int k = threadIdx.x; k < SMCB_NCZ; k += blockDim.x) zloc[k] = 0u;
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        // uint8_t i = floor((x/L+.5)*Ncx) ...  (SMC.c:917-919; the uint8_t wrap is kept)
        const int i = (int)floor((P[n] / cp.L + .5) * SMCB_NCX) & 0xff;
        const int j = (int)floor((P[Npad + n] / cp.L + .5) * SMCB_NCX) & 0xff;
        const int k = (int)floor((P[2 * Npad + n] / cp.Lz + .5) * SMCB_NCZ) & 0xff;
        const int v = i * SMCB_NCX * SMCB_NCZ + j * SMCB_NCZ + k;
        if (v < nvox) {
            atomicAdd(D + v, 1ull);
            int *rb = g.rbin + (size_t)chain * N + n;
            if (*rb != v) { atomicAdd(Mu + v, 1ull); *rb = v; }
        }
        if (k < SMCB_NCZ) atomicAdd(zloc + k, 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SMCB_NCZ; k += blockDim.x)
        if (zloc[k]) atomicAdd(zprof + k, (unsigned long long)zloc[k]);
    if (threadIdx.x == 0) {
        const double *t = g.totals + (size_t)chain * kTot;
        const double E = t[0] + t[1];
        const double vol3 = 3 * cp.L * cp.L * cp.Lz;
        const double Pv = -t[2] / vol3 + (-t[g.wall_virial_intended ? 4 : 3] / vol3);    // pressure() + wallsPressure()  SMC.c:140
        double *m = g.chain_mom + (size_t)chain * 5;
        const long long tri = d.ntri[chain];
        m[0] = E; m[1] = E * E; m[2] = Pv; m[3] = Pv * Pv;
        m[4] = tri > 0 ? (double)d.nacc[chain] / (double)tri : 0.0;
        const double epp = E / N;
        int bin = (int)floor((epp - g.e_lo) / (g.e_hi - g.e_lo) * g.nebins);
        bin = bin < 0 ? 0 : (bin >= g.nebins ? g.nebins - 1 : bin);
        atomicAdd(ehist + bin, 1ull);
        atomicAdd(nsamp, 1ull);
    }
}

// One block per observable group: the chains of the group are visited in chain order (thread t takes chains
// t, t + blockDim, ...; then a fixed-order block sum), and the single writer adds the totals to the group's moments.
__global__ void k_gather_moments(DevChains d, GatherArgs g)
{
    const int grp = blockIdx.x;
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        if ((int)chain_params(d, c).group != grp) continue;
        const double *m = g.chain_mom + (size_t)c * 5;
#pragma unroll
        for (int k = 0; k < 5; k++) v[k] += m[k];
    }
    __shared__ double scratch[8 * 32];
    block_sum<5>(v, scratch);
    if (threadIdx.x == 0) {
        double *m = g.moments + (size_t)grp * g.f64_per_group;
#pragma unroll
        for (int k = 0; k < 5; k++) m[k] += v[k];
    }
}

// ================================================================ helpers ===
// AoS [C][3N] (the reference's layout, SMC.h:84) <-> SoA [C][3][Npad]
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, int C, int N, int Npad, int ncomp)
{
    const size_t total = (size_t)C * Npad;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const size_t c = q / Npad;
        const int j = (int)(q - c * Npad);
        for (int k = 0; k < ncomp; k++)
            soa[(c * ncomp + k) * Npad + j] = (j < N) ? aos[(c * N + j) * ncomp + k] : 0.0;
    }
}

__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, int C, int N, int Npad, int ncomp)
{
    const size_t total = (size_t)C * N;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const size_t c = q / N;
        const int j = (int)(q - c * N);
        for (int k = 0; k < ncomp; k++)
            aos[(c * N + j) * ncomp + k] = soa[(c * ncomp + k) * Npad + j];
    }
}

// E[c] = U_lj + U_wall of the last evaluation (SMC.c:48: E[0] = energy(R) + wallsEnergy(R)), on the device
__global__ void k_energy_from_totals(const double *__restrict__ totals, double *__restrict__ E, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) E[c] = totals[(size_t)c * kTot] + totals[(size_t)c * kTot + 1];
}

// Extent of every chain's configuration in box units, for the FP32 screen's error bound (make_screen): one warp per
// chain.  Coordinates beyond 2^20 boxes (or NaN) cannot be screened in single precision at all: flagged.
__global__ void k_chain_extent(const double *__restrict__ pos, const smcb_chain_params *__restrict__ params, int nparams,
                               int C, int N, int Npad, float *__restrict__ extent, int *__restrict__ flag)
{
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chain >= C) return;
    const double invL = 1.0 / params[nparams == 1 ? 0 : chain].L;
    const double *P = pos + (size_t)chain * 3 * Npad;
    float axy = 0.f, az = 0.f;
    bool bad = false;
    for (int j = lane; j < N; j += 32) {
        const double x = fabs(P[j]) * invL, y = fabs(P[Npad + j]) * invL, z = fabs(P[2 * Npad + j]) * invL;
        bad |= !(x < 1048576.0) || !(y < 1048576.0) || !(z < 1048576.0);
        axy = fmaxf(axy, (float)fmax(x, y));
        az = fmaxf(az, (float)z);
    }
    for (int o = 16; o > 0; o >>= 1) {
        axy = fmaxf(axy, __shfl_xor_sync(FULL, axy, o));
        az = fmaxf(az, __shfl_xor_sync(FULL, az, o));
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(flag, 1);
    if (lane == 0) { extent[2 * chain] = axy * 1.0000002f; extent[2 * chain + 1] = az * 1.0000002f; }   // round up
}

// step-size control (pre-production only): A[c] *= exp(gain * (acceptance[c] - target)), counters cleared
__global__ void k_adapt_step(smcb_chain_params *params, long long *nacc, long long *ntri, int C, double target, double gain,
                             double a_min, double a_max)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long t = ntri[c];
    if (t > 0) {
        const double acc = (double)nacc[c] / (double)t;
        // nothing / everything accepted says only "far off": move by a decade / a factor of four instead of the gain's e-fold
        double A = params[c].A * (nacc[c] == 0 ? 0.1 : (nacc[c] == t ? 4.0 : exp(gain * (acc - target))));
        A = A < a_min ? a_min : (A > a_max ? a_max : A);
        params[c].A = A;
    }
    nacc[c] = 0; ntri[c] = 0;
}

// FP64 FMA peak: 8 independent DFMA chains per thread, no memory traffic
__global__ void k_dfma_peak(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-6;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;      // never true; keeps the chains alive
}
#endif  // SMCB_MISC_KERNELS

}  // namespace smcb
