// allparticle_fast.cuh — the FAST all-particle Smart-MC step (north-star kernel B).
//
// Every molecule of a chain is displaced at once, d_i = F_i A/T + xi_i, energy and forces are
// recomputed at the proposal with an O(N^2) pair kernel and ONE Metropolis-Hastings test per chain
// decides (acceptance expression of SMC.c:326-329 summed over the molecules; the reference's own
// all-particle attempt, markovProbability, is dead code, SMC.c:354-402).  Forces and energy of the
// current state are carried from step to step, so a step costs one force evaluation.
//
// Organisation for B200:
//  * CL thread blocks per chain (CL = 1, or a thread-block CLUSTER of CL = 2/4/8 for large N, so
//    that 32 chains of N = 4096 still fill 148 SMs).  Block `part` owns molecules i = part, part+CL...
//    strided by thread; every block stages ALL proposal positions of the chain in its shared memory
//    (the j side of the pair loop), recomputing the proposals redundantly - they cost O(N), the pair
//    loop O(N^2/CL).
//  * the pair loop is a packed single-precision SCREEN (box units, minimum image by the 1.5*2^23
//    trick, add/sub/mul/fma.f32x2 on two j per instruction; see sweep_cached.cuh for the error bound
//    that makes it a guaranteed superset) that collects a 32-bit hit mask per chunk of 32 j; the rare
//    hits are then evaluated in double precision from the exact positions.  A pair outside the cutoff
//    contributes exactly 0, so energies and forces are those of the all-FP64 loop, with the same
//    j-ascending summation order; the FP64 pipe (1 warp instruction / 2 cycles on B200) is left to
//    the pairs that matter.
//  * the three MH sums are reduced per block, then across the cluster through distributed shared
//    memory in rank order (deterministic); every block takes the same decision from the same
//    Philox number.
#pragma once
#include <cooperative_groups.h>

namespace smcb {
namespace cg = cooperative_groups;

struct StepSmem {
    double *x, *y, *z;        // proposal positions, exact                    [3][Npad]
    float *fx, *fy, *fz;      // the same in box units, screen precision      [3][Npad]
    double *scratch;          // block_sum scratch                            [8*32]
    double *part;             // this block's partial MH sums (read by the cluster peers) [4]
    // half-shell extras (HALF kernels only): the screen-precision coordinates once more as CIRCULAR arrays
    // (index m holds particle m mod N) in two copies, B shifted by one (B[m] = A[m+1]), so that the pair of
    // partners (i+1+2k, i+2+2k) of any i is one aligned float2; and the hit words of phase 1
    float *cbase;                   // [3 components][2 copies][NE]
    unsigned *fw, *bw;              // forward / backward hit words [N][NW]
    int NW, NE;
    __device__ __forceinline__ float *circ(int comp, int copy) const { return cbase + (2 * comp + copy) * NE; }
    static __host__ __device__ int half_words(int N) { return (N / 2 + 31) / 32; }
    static __host__ __device__ int half_ext(int N) { return (N + 32 * half_words(N) + 3) & ~1; }
    __device__ __forceinline__ void carve(double *base, int Npad, int N = 0, bool half = false)
    {
        x = base; y = x + Npad; z = y + Npad;
        scratch = z + Npad;
        part = scratch + 8 * 32;
        fx = reinterpret_cast<float *>(part + 4); fy = fx + Npad; fz = fy + Npad;
        NW = NE = 0;
        if (half) {
            NW = half_words(N); NE = half_ext(N);
            cbase = fz + Npad;
            fw = reinterpret_cast<unsigned *>(cbase + 6 * NE);
            bw = fw + (size_t)N * NW;
        }
    }
    static __host__ __device__ size_t bytes(int Npad, int N = 0, bool half = false)
    {
        size_t b = (size_t)(3 * Npad + 8 * 32 + 4) * sizeof(double) + (size_t)3 * Npad * sizeof(float);
        if (half) b += (size_t)6 * half_ext(N) * sizeof(float) + (size_t)2 * N * half_words(N) * sizeof(unsigned);
        return b;
    }
};

// LJ energy (already *4), force and in-cutoff count of molecule i at (px,py,pz) against the staged
// configuration; `self` is i's own index in that configuration (skipped)
// Every lane of the warp must call the pair-loop helpers below together (`act` = this lane has a molecule): after
// each divergent hit loop the warp is re-converged with __syncwarp(), otherwise the lanes that leave a hit loop
// early run ahead into the next chunk's screen on their own and the screens execute with a fraction of the warp
// (measured in the condensed phase: 8.5 active threads per instruction, 3 x the time).
template <bool PZ, bool VIR = false>
__device__ __forceinline__ void particle_vs_staged(const Box &b, const ScreenConsts &sc, const StepSmem &s, int N, int Npad,
                                                   bool act, int self, double px, double py, double pz,
                                                   double &e_lj, double &fx, double &fy, double &fz, unsigned &cnt,
                                                   double *vir = nullptr)
{
    const float qx = (float)(px * b.invL), qy = (float)(py * b.invL), qz = (float)(pz * b.invL);
    const float2 ax = make_float2(qx, qx), ay = make_float2(qy, qy), az = make_float2(qz, qz);
    const float2 MG = make_float2(12582912.f, 12582912.f);
    const float2 *X2 = reinterpret_cast<const float2 *>(s.fx), *Y2 = reinterpret_cast<const float2 *>(s.fy),
                 *Z2 = reinterpret_cast<const float2 *>(s.fz);
    double e = 0.0, v = 0.0;
    fx = fy = fz = 0.0;
    for (int c0 = 0; c0 < Npad; c0 += 32) {
        unsigned hits = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int j2 = (c0 >> 1) + k;
            float2 sx = sub2(ax, X2[j2]);
            sx = sub2(sx, sub2(add2(sx, MG), MG));
            float2 sy = sub2(ay, Y2[j2]);
            sy = sub2(sy, sub2(add2(sy, MG), MG));
            float2 sz = sub2(az, Z2[j2]);
            if (PZ) {
                const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
                sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
            }
            const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
            if (r2.x < sc.rc2s) hits |= 1u << (2 * k);
            if (r2.y < sc.rc2s) hits |= 2u << (2 * k);
        }
        if ((self >> 5) == (c0 >> 5)) hits &= ~(1u << (self & 31));
        if (!act) hits = 0u;
        while (hits) {                                   // rare in the gas; j ascending like the FP64 loop
            const int j = c0 + __ffs(hits) - 1;
            hits &= hits - 1;
            double et, gx, gy, gz;
            if (j < N && pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, gx, gy, gz)) {
                e += et; fx += gx; fy += gy; fz += gz;
                cnt++;
                if (VIR) {                               // pressure()'s pair term 24/r^6 - 48/r^12 (SMC.c:712-714)
                    double dx, dy, dz;
                    v += virial_term<false>(pair_sep<false>(b, px, py, pz, s.x[j], s.y[j], s.z[j], dx, dy, dz));
                }
            }
        }
        __syncwarp();
    }
    e_lj = 4.0 * e;
    if (VIR) *vir = v;
}

// one molecule against the surface, thread-serial: flat wall always (no cutoff, SMC.c:740-741), the M*M
// sites only when the molecule is within the cutoff of the wall plane (none can be in range otherwise).
// Returns the energy WITHOUT the final *4 and ADDS the force, like wall_particle.
__device__ __forceinline__ double wall_point_fast(const Box &b, const double *__restrict__ W, double px, double py, double pz,
                                                  double &fx, double &fy, double &fz)
{
    double e = 0.0;
    const double dzw = wall_dz<false>(b, pz);
    add_zwall(b, dzw, e, fz);
    if (dzw * dzw < b.rc2) {
        const double dw = b.L / b.M;
        for (int i = 0; i < b.M; i++)
            for (int j = 0; j < b.M; j++) {
                const int m = j + i * b.M;
                const double dx = min_image<false>(px - i * dw, b.L, b.invL);
                const double dy = min_image<false>(py - j * dw, b.L, b.invL);
                const double r2 = fma(dzw, dzw, fma(dy, dy, dx * dx));
                if (r2 < b.rc2) {
                    const double i2 = fast_rcp(r2);
                    const double i6 = i2 * i2 * i2;
                    const double a6 = W[2 * m] * i6;
                    e += fma(a6, i6, -W[2 * m + 1] * i6);
                    const double g = i2 * i6 * fma(48.0, a6, -24.0 * W[2 * m + 1]);
                    fx = fma(g, dx, fx);
                    fy = fma(g, dy, fy);
                    fz = fma(g, dzw, fz);
                }
            }
    }
    return e;
}

// ---- half-shell screen (Newton's third law for the screen) -------------------------------------------------
// Every unordered pair is screened ONCE: molecule i looks at the partners i+1 .. i+H (indices mod N, H = N/2;
// for even N the offset N/2 only from the lower half), which halves the O(N^2) part of a step.  A hit at offset
// d is recorded for both members: bit d-1 of fw[i] and, with a shared-memory atomicOr (order-independent), bit
// d-1 of bw[i+d].  Phase 2 then lets every molecule evaluate ALL its partners itself, forward offsets ascending,
// then backward offsets ascending - a fixed order, no floating-point atomics, and the pair terms seen from both
// members are exact negatives of each other.
template <bool PZ>
__device__ __forceinline__ void half_shell_screen(const ScreenConsts &sc, const StepSmem &s, int N, bool act, int i, float qx, float qy, float qz)
{
    const int H = N / 2, start = i + 1, odd = start & 1;
    const int Hi = (!(N & 1) && i >= H) ? H - 1 : H;          // valid offsets of this molecule
    const float2 *X2 = reinterpret_cast<const float2 *>(s.circ(0, odd)) + (start >> 1);
    const float2 *Y2 = reinterpret_cast<const float2 *>(s.circ(1, odd)) + (start >> 1);
    const float2 *Z2 = reinterpret_cast<const float2 *>(s.circ(2, odd)) + (start >> 1);
    const float2 ax = make_float2(qx, qx), ay = make_float2(qy, qy), az = make_float2(qz, qz);
    const float2 MG = make_float2(12582912.f, 12582912.f);
    for (int c = 0; c < s.NW; c++) {
        unsigned hits = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int j2 = c * 16 + k;
            float2 sx = sub2(ax, X2[j2]);
            sx = sub2(sx, sub2(add2(sx, MG), MG));
            float2 sy = sub2(ay, Y2[j2]);
            sy = sub2(sy, sub2(add2(sy, MG), MG));
            float2 sz = sub2(az, Z2[j2]);
            if (PZ) {
                const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
                sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
            }
            const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
            if (r2.x < sc.rc2s) hits |= 1u << (2 * k);
            if (r2.y < sc.rc2s) hits |= 2u << (2 * k);
        }
        const int rem = Hi - 32 * c;                          // offsets 32c+1 .. 32c+32 that exist
        if (rem < 32) hits &= rem <= 0 ? 0u : ((1u << rem) - 1u);
        if (!act) hits = 0u;
        if (act) s.fw[i * s.NW + c] = hits;
        while (hits) {
            const int bit = __ffs(hits) - 1;
            hits &= hits - 1;
            int j = i + 32 * c + bit + 1;
            if (j >= N) j -= N;
            atomicOr(s.bw + j * s.NW + c, 1u << bit);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void half_shell_exact(const Box &b, const StepSmem &s, int N, bool act, int i, double px, double py, double pz,
                                                 double &e_lj, double &fx, double &fy, double &fz, unsigned &cnt)
{
    double e = 0.0;
    fx = fy = fz = 0.0;
    for (int dir = 0; dir < 2; dir++) {
        const unsigned *words = (dir == 0 ? s.fw : s.bw) + i * s.NW;
        for (int c = 0; c < s.NW; c++) {
            unsigned w = act ? words[c] : 0u;
            while (w) {                                   // two partners per iteration, ascending offsets
                const int d1 = 32 * c + __ffs(w);
                w &= w - 1;
                const bool two = w != 0u;
                const int d2 = two ? 32 * c + __ffs(w) : d1;
                w &= w - 1;                               // (0 stays 0)
                int j1 = dir == 0 ? i + d1 : i - d1, j2 = dir == 0 ? i + d2 : i - d2;
                if (j1 >= N) j1 -= N;
                if (j1 < 0) j1 += N;
                if (j2 >= N) j2 -= N;
                if (j2 < 0) j2 += N;
                double e1, x1, y1, z1, e2, x2, y2, z2;
                const bool in1 = pair_terms_nb(b, px, py, pz, s.x[j1], s.y[j1], s.z[j1], e1, x1, y1, z1);
                const bool in2 = pair_terms_nb(b, px, py, pz, s.x[j2], s.y[j2], s.z[j2], e2, x2, y2, z2) && two;
                e += e1; fx += x1; fy += y1; fz += z1;
                if (in2) { e += e2; fx += x2; fy += y2; fz += z2; }
                cnt += (in1 ? 1u : 0u) + (in2 ? 1u : 0u);
            }
            __syncwarp();
        }
    }
    e_lj = 4.0 * e;
}

// FED: host-fed noise (parity); PZ: bulk mode (z periodic); CL: blocks per chain (cluster size);
// HALF: half-shell screen (one block per chain, N <= 512)
template <bool FED, bool PZ, int CL, bool HALF>
__device__ __forceinline__ void allparticle_fast_body(const DevChains &d, const StepArgs &a)
{
    static_assert(!HALF || CL == 1, "the half-shell screen keeps a chain in one block");
    const int chain = blockIdx.x / CL, part = blockIdx.x % CL;
    const int N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    extern __shared__ double sm[];
    StepSmem s;
    s.carve(sm, Npad, N, HALF);
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const ScreenConsts sc = make_screen(b, d.extent ? d.extent + 2 * chain : nullptr);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    double *Fc = a.F + (size_t)chain * 3 * Npad;
    double *Fn = a.Fn + (size_t)chain * 3 * Npad;
    double *DL = a.dl + (size_t)chain * 3 * Npad;

    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    unsigned cnt = 0;
    double U = d.E[chain];
    long long nacc = 0;

    auto cluster_barrier = [&]() {
        if (CL > 1) cg::this_cluster().sync();           // release/acquire at cluster scope: orders global and shared writes
        else __syncthreads();
    };
    // stage one configuration: exact + box-unit single precision; pad slots far away
    auto stage = [&](int j, double X, double Y, double Z) {
        s.x[j] = X; s.y[j] = Y; s.z[j] = Z;
        const float fx = (float)(X * b.invL), fy = (float)(Y * b.invL), fz = (float)(Z * b.invL);
        s.fx[j] = fx; s.fy[j] = fy; s.fz[j] = fz;
        if (HALF) {            // the circular copies (A[m] = particle m mod N, B[m] = A[m+1]) and cleared backward words
            for (int m = j; m < s.NE; m += N) { s.circ(0, 0)[m] = fx; s.circ(1, 0)[m] = fy; s.circ(2, 0)[m] = fz; }
            for (int m = (j == 0 ? N - 1 : j - 1); m < s.NE; m += N) { s.circ(0, 1)[m] = fx; s.circ(1, 1)[m] = fy; s.circ(2, 1)[m] = fz; }
            for (int c = 0; c < s.NW; c++) s.bw[j * s.NW + c] = 0u;
        }
    };
    auto stage_pad = [&]() {
        for (int j = N + tid; j < Npad; j += T_) { s.x[j] = 0.0; s.y[j] = 0.0; s.z[j] = 0.0; s.fx[j] = 0.f; s.fy[j] = 0.f; s.fz[j] = 3.0e18f; }
    };
    // forces / energies of this block's molecules at the staged configuration; returns the MH partial sums
    auto evaluate_owned = [&](double *Fout, const double *Fold, bool with_mh, double (&t)[3]) {
        t[0] = t[1] = t[2] = 0.0;
        if (HALF) {            // phase 1 (the caller's barrier after staging covers the circular copies and bw = 0)
            for (int i0 = 0; i0 < N; i0 += T_) {           // uniform trip count: the helpers re-converge the warp
                const int i = i0 + tid;
                const bool act = i < N;
                const int ic = act ? i : 0;
                half_shell_screen<PZ>(sc, s, N, act, ic, s.fx[ic], s.fy[ic], s.fz[ic]);
            }
            __syncthreads();
        }
        for (int i0 = 0; i0 < N; i0 += CL * T_) {
            const int i = i0 + part + CL * tid;
            const bool act = i < N;
            const int ic = act ? i : 0;
            const double px = s.x[ic], py = s.y[ic], pz = s.z[ic];
            double e_lj, fx, fy, fz;
            if (HALF) half_shell_exact(b, s, N, act, ic, px, py, pz, e_lj, fx, fy, fz, cnt);
            else particle_vs_staged<PZ>(b, sc, s, N, Npad, act, ic, px, py, pz, e_lj, fx, fy, fz, cnt);
            if (!act) continue;
            double e_wall = 0.0;
            if (b.wall) {
                double wx = 0.0, wy = 0.0, wz = 0.0;
                e_wall = wall_point_fast(b, W, px, py, pz, wx, wy, wz) * 4;
                fx += wx; fy += wy; fz += wz;
            }
            Fout[i] = fx; Fout[Npad + i] = fy; Fout[2 * Npad + i] = fz;
            t[0] += 0.5 * e_lj + e_wall;
            if (with_mh) {
                const double ox = Fold[i], oy = Fold[Npad + i], oz = Fold[2 * Npad + i];
                t[1] += DL[i] * (fx + ox) + DL[Npad + i] * (fy + oy) + DL[2 * Npad + i] * (fz + oz);
                t[2] += (fx * fx - ox * ox) + (fy * fy - oy * oy) + (fz * fz - oz * oz);
            }
        }
        block_sum<3>(t, s.scratch);
        if (CL > 1) {                                     // cluster-wide sums through distributed shared memory, rank order
            cg::cluster_group cl = cg::this_cluster();
            if (tid == 0) { s.part[0] = t[0]; s.part[1] = t[1]; s.part[2] = t[2]; }
            cl.sync();
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            for (int r = 0; r < CL; r++) {
                const double *peer = cl.map_shared_rank(s.part, r);
                r0 += peer[0]; r1 += peer[1]; r2 += peer[2];
            }
            cl.sync();                                    // peers are done reading before `part` is reused
            t[0] = r0; t[1] = r1; t[2] = r2;
        }
    };

    stage_pad();
    if (a.refresh) {           // bring F and U in line with the positions
        for (int j = tid; j < N; j += T_) stage(j, P[j], P[Npad + j], P[2 * Npad + j]);
        __syncthreads();
        double t[3];
        evaluate_owned(Fc, Fc, false, t);
        U = t[0];
        cnt = 0;
        cluster_barrier();     // every block's share of Fc is visible before the first proposal reads it
    }

    for (int st = 0; st < a.nsteps; st++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)st;
        const size_t sci = (size_t)st * d.C + chain;
        // ---- proposal: d_j = F_j A/T + xi_j ; r' = wrap(r + d), all molecules, staged for the pair loop
        for (int j = tid; j < N; j += T_) {
            double g0, g1, g2;
            if (FED) {
                const double *xi = a.xi + sci * 3 * N;
                g0 = xi[3 * j]; g1 = xi[3 * j + 1]; g2 = xi[3 * j + 2];
            } else {
                rng_particle_gauss_f32(id, step, (uint32_t)j, g0, g1, g2);
                g0 *= sigma; g1 *= sigma; g2 *= sigma;
            }
            const double dX = fma(Fc[j], AoT, g0), dY = fma(Fc[Npad + j], AoT, g1), dZ = fma(Fc[2 * Npad + j], AoT, g2);
            if (j % CL == part) { DL[j] = dX; DL[Npad + j] = dY; DL[2 * Npad + j] = dZ; }      // owner keeps the displacement
            double qx = P[j] + dX, qy = P[Npad + j] + dY, qz = P[2 * Npad + j] + dZ;
            qx = min_image<false>(qx, b.L, b.invL);
            qy = min_image<false>(qy, b.L, b.invL);
            if (PZ) qz = min_image<false>(qz, b.Lz, b.invLz);
            stage(j, qx, qy, qz);
        }
        __syncthreads();
        // ---- forces and energy at the proposal, MH sums
        double t[3];                       // U', sum d.(F'+F), sum |F'|^2-|F|^2
        evaluate_owned(Fn, Fc, true, t);
        const double lnap = -((t[0] - U) + t[1] / 2.0 + t[2] * b.A / (4.0 * b.T)) / b.T;
        // every thread of every block of the chain holds the same sums and draws the same Philox number: the
        // decision needs no broadcast and no barrier
        double uu;
        if (FED) uu = a.u[sci];
        else { uint32_t o; rng_step_scalars(id, step, o, uu); }
        const bool acc = uu < exp(lnap);
        if (tid == 0 && part == 0) {
            if (a.lnap) a.lnap[sci] = lnap;
            if (a.accepted) a.accepted[sci] = (unsigned char)acc;
        }
        if (acc) {
            for (int i = part + CL * tid; i < N; i += CL * T_) { P[i] = s.x[i]; P[Npad + i] = s.y[i]; P[2 * Npad + i] = s.z[i]; }
            double *tp = Fc; Fc = Fn; Fn = tp;
            U = t[0];
            nacc++;
        }
        cluster_barrier();                 // positions / forces of this step are visible to the next proposal
    }

    double *Fcanon = a.F + (size_t)chain * 3 * Npad;
    if (Fc != Fcanon)
        for (int i = part + CL * tid; i < N; i += CL * T_) { Fcanon[i] = Fc[i]; Fcanon[Npad + i] = Fc[Npad + i]; Fcanon[2 * Npad + i] = Fc[2 * Npad + i]; }
    double c[1] = {(double)cnt};
    block_sum<1>(c, s.scratch);
    if (tid == 0) {
        if (part == 0) {
            d.E[chain] = U;
            d.nacc[chain] += nacc;
            d.ntri[chain] += a.nsteps;
            if (d.pair_counts) {
                atomicAdd(d.pair_counts, (unsigned long long)a.nsteps * (unsigned long long)N * (N - 1));
                // screened pair tests: every ordered pair, or every unordered pair once in the half-shell variant
                atomicAdd(d.pair_counts + 2, (unsigned long long)(a.nsteps + (a.refresh ? 1 : 0)) * (unsigned long long)N * (N - 1) / (HALF ? 2ull : 1ull));
            }
        }
        if (d.pair_counts) atomicAdd(d.pair_counts + 1, (unsigned long long)c[0]);
    }
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP): one contiguous tile global -> shared, completion on an mbarrier.
// A chain's positions are one contiguous [3][Npad] block in HBM (SoA, Npad a multiple of 32: 16-byte aligned and
// sized), exactly the layout of StepSmem::x/y/z, so the whole j-side tile arrives with ONE asynchronous copy issued
// by one thread while the others fill the pad slots; nobody spends LSU instructions or registers on it.
__device__ __forceinline__ void tile_load_begin(double *smem_dst, const double *gsrc, uint32_t bytes, uint64_t *bar)
{
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst_a = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_a), "l"(gsrc), "r"(bytes), "r"(bar_a) : "memory");
}

__device__ __forceinline__ void tile_load_wait(uint64_t *bar)
{
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
}

// ---- FAST static evaluation (rows a2-a10 of the survey) with the same screened pair loop ----------------
// `parts` blocks per chain (plain grid, no cluster): block `part` owns molecules part, part+parts, ...; every
// block stages the whole chain.  Chain totals: each block leaves its partial sums in `partials`, and the LAST
// block of a chain to finish (atomic ticket) adds them in part order, so the totals do not depend on scheduling.
struct EvalFastArgs {
    int parts;
    double *partials;          // [C][parts][kTot]
    unsigned *tickets;         // [C], zeroed by the launcher's caller, left zero again by the kernel
};

template <bool PZ>
__device__ __forceinline__ void evaluate_fast_body(const DevChains &d, const EvalOut &o, const EvalFastArgs &ea)
{
    const int parts = ea.parts, chain = blockIdx.x / parts, part = blockIdx.x % parts;
    const int N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    extern __shared__ double sm[];
    StepSmem s;
    s.carve(sm, Npad);
    __shared__ unsigned s_last;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, 1.0);
    const ScreenConsts sc = make_screen(b, d.extent ? d.extent + 2 * chain : nullptr);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    const double *P = d.pos + (size_t)chain * 3 * Npad;
    // the chain's [3][Npad] position block -> s.x/s.y/s.z in one bulk-async copy (pad slots hold 0 in HBM)
    __shared__ uint64_t s_bar;
    if (tid == 0) tile_load_begin(s.x, P, (uint32_t)(3 * Npad * sizeof(double)), &s_bar);
    __syncthreads();                                      // the barrier is initialised before anyone polls it
    tile_load_wait(&s_bar);
    for (int j = tid; j < Npad; j += T_) {                // screen-precision copy in box units; pad slots far away
        s.fx[j] = (float)(s.x[j] * b.invL); s.fy[j] = (float)(s.y[j] * b.invL);
        s.fz[j] = j < N ? (float)(s.z[j] * b.invL) : 3.0e18f;
    }
    __syncthreads();
    double tot[kTot] = {0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned cnt = 0;
    for (int i0 = 0; i0 < N; i0 += parts * T_) {            // uniform trip count: the pair loop re-converges the warp
        const int i = i0 + part + parts * tid;
        const bool act = i < N;
        const int ic = act ? i : 0;
        const double px = s.x[ic], py = s.y[ic], pz = s.z[ic];
        double e_lj, fx, fy, fz, vir;
        particle_vs_staged<PZ, true>(b, sc, s, N, Npad, act, ic, px, py, pz, e_lj, fx, fy, fz, cnt, &vir);
        if (!act) continue;
        double e_wall = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
        if (b.wall) {
            e_wall = wall_point_fast(b, W, px, py, pz, wx, wy, wz) * 4;
            tot[3] += wall_virial_ref<false>(b, W, px, py, pz);
            tot[4] += wall_virial_intended<false>(b, W, px, py, pz);
        }
        const size_t q = (size_t)chain * Npad + i, q3 = (size_t)chain * 3 * Npad + i;
        if (o.e_lj) o.e_lj[q] = e_lj;
        if (o.e_wall) o.e_wall[q] = e_wall;
        if (o.f_lj) { o.f_lj[q3] = fx; o.f_lj[q3 + Npad] = fy; o.f_lj[q3 + 2 * Npad] = fz; }
        if (o.f_wall) { o.f_wall[q3] = wx; o.f_wall[q3 + Npad] = wy; o.f_wall[q3 + 2 * Npad] = wz; }
        tot[0] += 0.5 * e_lj;
        tot[1] += e_wall;
        tot[2] += 0.5 * vir;
    }
    block_sum<kTot>(tot, s.scratch);
    if (!o.totals) return;
    double *mine = ea.partials + ((size_t)chain * parts + part) * kTot;
    if (tid == 0) {
        for (int k = 0; k < kTot; k++) mine[k] = tot[k];
        __threadfence();
        s_last = atomicAdd(ea.tickets + chain, 1u) == (unsigned)(parts - 1);
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        const double *pp = ea.partials + (size_t)chain * parts * kTot;
        double r[kTot] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int p = 0; p < parts; p++)
            for (int k = 0; k < kTot; k++) r[k] += pp[kTot * p + k];
        double *t = o.totals + (size_t)chain * kTot;
        for (int k = 0; k < kTot; k++) t[k] = r[k];
        ea.tickets[chain] = 0;
    }
}

#ifdef SMCB_MISC_KERNELS      // not a template: compiled once, in kernels_fast.cu
__global__ void __launch_bounds__(512) k_evaluate_fast(DevChains d, EvalOut o, EvalFastArgs ea)
{
    if (chain_params(d, blockIdx.x / ea.parts).flags & SMCB_PERIODIC_Z) evaluate_fast_body<true>(d, o, ea);
    else evaluate_fast_body<false>(d, o, ea);
}
#endif

template <bool FED, int CL, bool HALF>
__global__ void __launch_bounds__(512) k_allparticle_fast(DevChains d, StepArgs a)
{
    if (chain_params(d, blockIdx.x / CL).flags & SMCB_PERIODIC_Z) allparticle_fast_body<FED, true, CL, HALF>(d, a);
    else allparticle_fast_body<FED, false, CL, HALF>(d, a);
}

}  // namespace smcb
