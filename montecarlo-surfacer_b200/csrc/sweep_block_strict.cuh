// sweep_block_strict.cuh — the bit-exact sweep (oneParticleMoves, SMC.c:278-351) for N > 512: one block per chain.
//
// k_sweep<K, STRICT> keeps a chain's molecules in the registers of one warp, which stops at N = 512.  Beyond that the
// drop-in's oneParticleMoves used to fall back to FAST arithmetic, so "the same trajectory as the reference, bit for
// bit" ended at 512 molecules.  This kernel keeps the reference's arithmetic (instantiated only in kernels_strict.cu,
// --fmad=false: no contraction, true divisions, the C expressions of SMC.c written as they read) and its SUMMATION
// ORDER at any N that fits shared memory:
//   phase A  all threads test their molecules l against the point (pair_sep<STRICT>); a warp's ballot IS the word of a
//            bitmap "l is inside the cutoff" over ascending l (warp w, iteration it -> molecules 32 (w + NW it) ...)
//   phase B  warp 0 walks the bitmap upwards; the lanes whose bit is set form the 12-6 terms of their molecule, and the
//            terms are added one by one in ascending l through shuffles, every lane keeping identical accumulators -
//            exactly energySingle's and forceSingle's loops over l (SMC.c:563-581, 597-617) restricted to the terms that
//            are not skipped; then the flat wall and the sites m ascending (SMC.c:735-761, 783-811).
// Two such evaluations per trial (old and proposed position), like the reference; no caches.  A parity path: one chain
// of N = 4096 runs ~25 sweeps/s (the reference on one host core: 2.4), the throughput kernels are the FAST ones.
#pragma once

namespace smcb {

struct StrictBlockSmem {
    double *x, *y, *z;          // [Npad]
    unsigned *bits;             // [Npad/32]
    double *res;                // [4] U, Fx, Fy, Fz of the last evaluation
    static __host__ __device__ size_t bytes(int Npad) { return (size_t)(3 * Npad + 8) * sizeof(double) + (size_t)(Npad / 32 + 4) * sizeof(unsigned); }
};

// energySingle + wallsEnergySingle and forceSingle + wallsForce of the molecule `self` placed at p (SMC.c:300-304, 319-321)
__device__ __forceinline__ void strict_block_eval(const Box &b, const double *__restrict__ W, const StrictBlockSmem &s, int N, int Npad,
                                                  int self, double px, double py, double pz, unsigned long long &cnt)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int nwords = Npad >> 5;
    for (int wi = warp; wi < nwords; wi += NW) {
        const int l = 32 * wi + lane;
        double dx, dy, dz;
        const double r2 = pair_sep<true>(b, px, py, pz, s.x[l], s.y[l], s.z[l], dx, dy, dz);
        const unsigned word = __ballot_sync(FULL, (r2 < b.rc2) && l < N && l != self);
        if (lane == 0) s.bits[wi] = word;
    }
    __syncthreads();
    if (warp == 0) {
        double V = 0.0, Fx = 0.0, Fy = 0.0, Fz = 0.0;
        for (int c0 = 0; c0 < nwords; c0 += 32) {
            const unsigned mine = (c0 + lane < nwords) ? s.bits[c0 + lane] : 0u;
            unsigned nz = __ballot_sync(FULL, mine != 0u);
            while (nz) {
                const int src = __ffs(nz) - 1;
                nz &= nz - 1;
                unsigned word = __shfl_sync(FULL, mine, src);
                const int l = 32 * (c0 + src) + lane;
                double et = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
                if ((word >> lane) & 1u) {
                    double dx, dy, dz, g;
                    const double r2 = pair_sep<true>(b, px, py, pz, s.x[l], s.y[l], s.z[l], dx, dy, dz);
                    lj_terms<true, true>(r2, 1.0, 1.0, et, g);
                    gx = g * dx; gy = g * dy; gz = g * dz;
                    cnt++;
                }
                while (word) {                            // ascending l: the reference's order
                    const int k = __ffs(word) - 1;
                    word &= word - 1;
                    V += __shfl_sync(FULL, et, k);
                    Fx += __shfl_sync(FULL, gx, k);
                    Fy += __shfl_sync(FULL, gy, k);
                    Fz += __shfl_sync(FULL, gz, k);
                }
            }
        }
        double Vw = 0.0;
        if (b.wall) {                                     // flat wall first, then sites m = i*M + j ascending
            const int MM = b.M * b.M;
            const double dw = b.L / b.M;
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
                        const int k = __ffs(mask) - 1;
                        mask &= mask - 1;
                        Vw += __shfl_sync(FULL, et, k);
                        Fx += __shfl_sync(FULL, gx, k);
                        Fy += __shfl_sync(FULL, gy, k);
                        Fz += __shfl_sync(FULL, gz, k);
                    }
                }
            }
        }
        if (lane == 0) { s.res[0] = V * 4 + Vw * 4; s.res[1] = Fx; s.res[2] = Fy; s.res[3] = Fz; }   // SMC.c:300
    }
    __syncthreads();
}

template <bool FED>
__global__ void k_sweep_block_strict(DevChains d, SweepArgs a)
{
    const int tid = threadIdx.x, T_ = blockDim.x, chain = blockIdx.x;
    const int N = d.N, Npad = d.Npad;
    extern __shared__ double sm[];
    StrictBlockSmem s;
    s.x = sm; s.y = s.x + Npad; s.z = s.y + Npad;
    s.res = s.z + Npad;
    s.bits = reinterpret_cast<unsigned *>(s.res + 8);
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    for (int j = tid; j < Npad; j += T_) {
        const bool in = j < N;
        s.x[j] = in ? P[j] : 0.0; s.y[j] = in ? P[Npad + j] : 0.0; s.z[j] = in ? P[2 * Npad + j] : 0.0;
    }
    __syncthreads();
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    double E = d.E[chain];
    long long nacc = 0;
    unsigned long long cnt = 0;
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};

    for (int sw = 0; sw < a.nsweeps; sw++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)sw;
        const size_t sc = (size_t)sw * d.C + chain;
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
        for (int nn = 0; nn < N; nn++) {
            int n = nn + off;                          // n = (nn+offset)%N  SMC.c:294
            if (n >= N) n -= N;
            double gx, gy, gz, uu;
            if (FED) {
                const double *dsp = a.displ + sc * 3 * N;
                gx = dsp[3 * n]; gy = dsp[3 * n + 1]; gz = dsp[3 * n + 2];
                uu = a.u[sc * N + nn];
            } else {                                   // every thread draws the same numbers (counter-based stream)
                rng_particle_gauss(id, step, (uint32_t)n, gx, gy, gz);
                gx *= sigma; gy *= sigma; gz *= sigma;
                uu = rng_particle_uniform(id, step, (uint32_t)n);
            }
            const double px = s.x[n], py = s.y[n], pz = s.z[n];
            strict_block_eval(b, W, s, N, Npad, n, px, py, pz, cnt);              // SMC.c:300-304
            const double Um = s.res[0], Fmx = s.res[1], Fmy = s.res[2], Fmz = s.res[3];
            const double dX = Fmx * b.A / b.T + gx, dY = Fmy * b.A / b.T + gy, dZ = Fmz * b.A / b.T + gz;   // SMC.c:307-309
            double qx = px + dX, qy = py + dY, qz = pz + dZ;                        // SMC.c:311-316
            qx = min_image<true>(qx, b.L, b.invL);
            qy = min_image<true>(qy, b.L, b.invL);
            if (b.pz) qz = min_image<true>(qz, b.Lz, b.invLz);
            strict_block_eval(b, W, s, N, Npad, n, qx, qy, qz, cnt);              // SMC.c:319-321
            const double Un = s.res[0], Fnx = s.res[1], Fny = s.res[2], Fnz = s.res[3];
            const double hx = Fnx - Fmx, hy = Fny - Fmy, hz = Fnz - Fmz;            // SMC.c:326-329
            const double dWk = (hx * hx + hy * hy + hz * hz + 2.0 * (hx * Fmx + hy * Fmy + hz * Fmz)) * b.A / (4.0 * b.T);
            const double ap = exp(-(Un - Um + (dX * (Fnx + Fmx) + dY * (Fny + Fmy) + dZ * (Fnz + Fmz)) / 2.0 + dWk) / b.T);
            const bool acc = uu < ap;                  // SMC.c:335
            if (acc) {
                if (tid == 0) { s.x[n] = qx; s.y[n] = qy; s.z[n] = qz; }
                E += Un - Um;                          // SMC.c:341
                nacc++;
            }
            if (FED && a.accepted != nullptr && tid == 0) a.accepted[sc * N + nn] = acc ? 1 : 0;
            __syncthreads();
        }
        if (a.trace_E != nullptr && tid == 0) { a.trace_E[sc] = E; a.trace_acc[sc] = (int)(nacc - nacc0); }
    }
    for (int j = tid; j < N; j += T_) { P[j] = s.x[j]; P[Npad + j] = s.y[j]; P[2 * Npad + j] = s.z[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);       // warp 0 counted, lane by lane
    if (tid == 0) {
        d.E[chain] = E;
        d.nacc[chain] += nacc;
        d.ntri[chain] += (long long)a.nsweeps * N;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
            atomicAdd(d.pair_counts + 1, cnt);
            atomicAdd(d.pair_counts + 2, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The same bit-exact sweep, batch-speculative (the organisation of sweep_block_spec.cuh): every warp of the block
// evaluates ONE trial of the next batch on its own - energySingle + wallsEnergySingle and forceSingle + wallsForce at
// the old position, the proposal, the same at the proposal, the acceptance (SMC.c:300-335) - against the positions at
// the start of the batch; the prefix of the batch that no earlier ACCEPTED trial can have influenced is committed in
// visiting order.  What makes it STRICT: the packed-FP32 screen only preselects (a superset of the partners; the
// reference's own test r2 < rc^2 in its own arithmetic decides), the 12-6 terms are the reference's expressions
// (--fmad=false), and they are added one by one in ascending partner index - screen iteration by screen iteration,
// lane by lane, through shuffles, every lane keeping identical accumulators.  A committed trial therefore produces the
// bits of k_sweep_block_strict and of the reference; only the time changes (a warp instead of a block per trial, three
// barriers per batch instead of four per trial).
template <bool PZ>
__device__ __forceinline__ void strict_warp_eval(const Box &b, const ScreenConsts &sc, const double *__restrict__ W, const BlockSpecSmem &s,
                                                 int N, int nit, int self, double px, double py, double pz, int lane,
                                                 double &U, double &Fx_, double &Fy_, double &Fz_, unsigned long long &cnt)
{
    const float qx[1] = {(float)(px * b.invL)}, qy[1] = {(float)(py * b.invL)}, qz[1] = {(float)(pz * b.invL)};
    unsigned long long hq[1][kBlockSpecWords / 2];
    const bool wx = !(fabsf(qx[0]) < sc.interior), wy = !(fabsf(qy[0]) < sc.interior);
    if (wx) {
        if (wy) block_spec_screen<PZ, true, true, 1>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, true, false, 1>(sc, s, nit, lane, qx, qy, qz, hq);
    } else {
        if (wy) block_spec_screen<PZ, false, true, 1>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, false, false, 1>(sc, s, nit, lane, qx, qy, qz, hq);
    }
    double V = 0.0, Fx = 0.0, Fy = 0.0, Fz = 0.0;
#pragma unroll
    for (int c = 0; c < kBlockSpecWords / 2; c++) {
        // which of the word's 32 screen iterations hold a candidate in ANY lane (two bits per iteration, first on top)
        const unsigned lo = __reduce_or_sync(FULL, (unsigned)hq[0][c]), hi = __reduce_or_sync(FULL, (unsigned)(hq[0][c] >> 32));
        unsigned long long any = ((unsigned long long)hi << 32) | lo;
        while (any) {
            const int il = __clzll((long long)any) >> 1;
            any &= ~(0xC000000000000000ull >> (2 * il));
            const unsigned two = (unsigned)(hq[0][c] >> (62 - 2 * il)) & 3u;
            const int j0 = 2 * (lane + 32 * (32 * c + il));            // this lane's molecules of the iteration: j0, j0 + 1
            double e0 = 0.0, x0 = 0.0, y0 = 0.0, z0 = 0.0, e1 = 0.0, x1 = 0.0, y1 = 0.0, z1 = 0.0;
            bool in0 = false, in1 = false;
            if ((two & 2u) && j0 != self && j0 < N) {
                double dx, dy, dz, g;
                const double r2 = pair_sep<true>(b, px, py, pz, s.x[j0], s.y[j0], s.z[j0], dx, dy, dz);
                if (r2 < b.rc2) { lj_terms<true, true>(r2, 1.0, 1.0, e0, g); x0 = g * dx; y0 = g * dy; z0 = g * dz; in0 = true; }
            }
            if ((two & 1u) && j0 + 1 != self && j0 + 1 < N) {
                double dx, dy, dz, g;
                const double r2 = pair_sep<true>(b, px, py, pz, s.x[j0 + 1], s.y[j0 + 1], s.z[j0 + 1], dx, dy, dz);
                if (r2 < b.rc2) { lj_terms<true, true>(r2, 1.0, 1.0, e1, g); x1 = g * dx; y1 = g * dy; z1 = g * dz; in1 = true; }
            }
            const unsigned m0 = __ballot_sync(FULL, in0), m1 = __ballot_sync(FULL, in1);
            unsigned m = m0 | m1;
            while (m) {                                            // ascending l = ascending lane, even molecule first
                const int k = __ffs(m) - 1;
                m &= m - 1;
                if ((m0 >> k) & 1u) {
                    V += __shfl_sync(FULL, e0, k); Fx += __shfl_sync(FULL, x0, k); Fy += __shfl_sync(FULL, y0, k); Fz += __shfl_sync(FULL, z0, k);
                }
                if ((m1 >> k) & 1u) {
                    V += __shfl_sync(FULL, e1, k); Fx += __shfl_sync(FULL, x1, k); Fy += __shfl_sync(FULL, y1, k); Fz += __shfl_sync(FULL, z1, k);
                }
            }
            cnt += (unsigned long long)(__popc(m0) + __popc(m1));
        }
    }
    double Vw = 0.0;
    if (b.wall) {                                                 // flat wall first, then sites m = i*M + j ascending
        const int MM = b.M * b.M;
        const double dw = b.L / b.M;
        const double dzw = wall_dz<true>(b, pz);
        double e0, g0;
        zwall_terms<true>(b, dzw, e0, g0);
        Vw += e0;
        Fz += g0 * dzw;
        for (int mb = 0; mb < MM; mb += 32) {
            const int m = mb + lane;
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
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    Vw += __shfl_sync(FULL, et, k);
                    Fx += __shfl_sync(FULL, gx, k);
                    Fy += __shfl_sync(FULL, gy, k);
                    Fz += __shfl_sync(FULL, gz, k);
                }
            }
        }
    }
    U = V * 4 + Vw * 4; Fx_ = Fx; Fy_ = Fy; Fz_ = Fz;             // SMC.c:300
}

template <bool FED, bool PZ>
__device__ __forceinline__ void sweep_block_strict_spec_body(const DevChains &d, const SweepArgs &a)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, NW = T_ >> 5;
    extern __shared__ double sm[];
    BlockSpecSmem s;
    s.carve(sm, Npad);
    const int NF = BlockSpecSmem::nf(Npad), nit = NF >> 6;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const ScreenConsts sc = make_screen(b, d.extent ? d.extent + 2 * chain : nullptr);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    for (int j = tid; j < NF; j += T_) {
        const bool in = j < N;
        const double X = in ? P[j] : 0.0, Y = in ? P[Npad + j] : 0.0, Z = in ? P[2 * Npad + j] : 0.0;
        if (j < Npad) { s.x[j] = X; s.y[j] = Y; s.z[j] = Z; }
        s.fx[j] = (float)(X * b.invL); s.fy[j] = (float)(Y * b.invL); s.fz[j] = in ? (float)(Z * b.invL) : 3.0e18f;
    }
    __syncthreads();
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    double E = d.E[chain];                           // kept by thread 0
    long long nacc = 0;
    unsigned long long cnt = 0, nscr = 0;            // thread 0
    const int BMAX = min(32, NW);

    for (int sw = 0; sw < a.nsweeps; sw++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)sw;
        const size_t sci = (size_t)sw * d.C + chain;
        const long long nacc0 = nacc;
        long long offset;                              // int offset = rand();  SMC.c:290
        if (FED) offset = a.offset[sci];
        else { uint32_t o; double unused; rng_step_scalars(id, step, o, unused); offset = o; }
        const int off = (int)(offset % N);
        int nn0 = 0;
        while (nn0 < N) {
            const int Bn = min(BMAX, N - nn0);
            bool acc = false;
            double qx = 0.0, qy = 0.0, qz = 0.0;
            float ox = 0.f, oy = 0.f, oz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
            int n = 0;
            if (warp < Bn) {
                const int nn = nn0 + warp;
                n = nn + off;                          // n = (nn+offset)%N  SMC.c:294
                if (n >= N) n -= N;
                double gx, gy, gz, uu;
                if (FED) {
                    const double *dsp = a.displ + sci * 3 * N;
                    gx = dsp[3 * n]; gy = dsp[3 * n + 1]; gz = dsp[3 * n + 2];
                    uu = a.u[sci * N + nn];
                } else {
                    rng_particle_gauss(id, step, (uint32_t)n, gx, gy, gz);
                    gx *= sigma; gy *= sigma; gz *= sigma;
                    uu = rng_particle_uniform(id, step, (uint32_t)n);
                }
                const double px = s.x[n], py = s.y[n], pz = s.z[n];
                double Um, Fmx, Fmy, Fmz, Un, Fnx, Fny, Fnz;
                unsigned long long nin = 0;
                strict_warp_eval<PZ>(b, sc, W, s, N, nit, n, px, py, pz, lane, Um, Fmx, Fmy, Fmz, nin);           // SMC.c:300-304
                const double dX = Fmx * b.A / b.T + gx, dY = Fmy * b.A / b.T + gy, dZ = Fmz * b.A / b.T + gz;   // SMC.c:307-309
                qx = px + dX; qy = py + dY; qz = pz + dZ;                                                       // SMC.c:311-316
                qx = min_image<true>(qx, b.L, b.invL);
                qy = min_image<true>(qy, b.L, b.invL);
                if (PZ) qz = min_image<true>(qz, b.Lz, b.invLz);
                strict_warp_eval<PZ>(b, sc, W, s, N, nit, n, qx, qy, qz, lane, Un, Fnx, Fny, Fnz, nin);           // SMC.c:319-321
                const double hx = Fnx - Fmx, hy = Fny - Fmy, hz = Fnz - Fmz;                                    // SMC.c:326-329
                const double dWk = (hx * hx + hy * hy + hz * hz + 2.0 * (hx * Fmx + hy * Fmy + hz * Fmz)) * b.A / (4.0 * b.T);
                const double ap = exp(-(Un - Um + (dX * (Fnx + Fmx) + dY * (Fny + Fmy) + dZ * (Fnz + Fmz)) / 2.0 + dWk) / b.T);
                acc = uu < ap;                         // SMC.c:335
                ox = s.fx[n]; oy = s.fy[n]; oz = s.fz[n];
                nx = (float)(qx * b.invL); ny = (float)(qy * b.invL); nz = (float)(qz * b.invL);
                bool hit = false;                      // earlier trials' molecules in range of my old or proposed position?
                if (lane < warp) {
                    int nm = nn0 + lane + off;
                    if (nm >= N) nm -= N;
                    const float mx = s.fx[nm], my = s.fy[nm], mz = s.fz[nm];
                    hit = screen_near<PZ>(sc, ox, oy, oz, mx, my, mz) || screen_near<PZ>(sc, nx, ny, nz, mx, my, mz);
                }
                const unsigned x1 = __ballot_sync(FULL, hit);
                if (lane == 0) {
                    s.pq[warp] = qx; s.pq[32 + warp] = qy; s.pq[64 + warp] = qz;
                    s.pf[warp] = nx; s.pf[32 + warp] = ny; s.pf[64 + warp] = nz;
                    s.pdU[warp] = Un - Um;
                    s.pacc[warp] = acc ? 1u : 0u; s.pin[warp] = (unsigned)nin; s.px[warp] = x1;
                }
            }
            __syncthreads();
            if (warp < Bn) {                           // ... and their proposals
                bool hit = false;
                if (lane < warp) {
                    const float mx = s.pf[lane], my = s.pf[32 + lane], mz = s.pf[64 + lane];
                    hit = screen_near<PZ>(sc, ox, oy, oz, mx, my, mz) || screen_near<PZ>(sc, nx, ny, nz, mx, my, mz);
                }
                const unsigned x2 = __ballot_sync(FULL, hit);
                if (lane == 0) s.px[warp] |= x2;
            }
            __syncthreads();
            const unsigned accmask = __ballot_sync(FULL, lane < Bn && s.pacc[lane] != 0u);
            const unsigned bad = __ballot_sync(FULL, lane < Bn && (s.px[lane] & accmask) != 0u);
            const int f = bad ? __ffs(bad) - 1 : Bn;   // the first trial with an accepted predecessor in range
            if (warp < f && acc && lane == 0) {
                s.x[n] = qx; s.y[n] = qy; s.z[n] = qz;
                s.fx[n] = nx; s.fy[n] = ny; s.fz[n] = nz;
            }
            if (tid == 0) {
                for (int w = 0; w < f; w++) {
                    cnt += s.pin[w];
                    const bool aw = (accmask >> w) & 1u;
                    if (aw) { E += s.pdU[w]; nacc++; }                 // SMC.c:341, in visiting order
                    if (FED && a.accepted != nullptr) a.accepted[sci * N + nn0 + w] = aw ? 1 : 0;
                }
                nscr += (unsigned long long)Bn * 2ull * (unsigned long long)(N - 1);
            }
            __syncthreads();
            nn0 += f;
        }
        if (a.trace_E != nullptr && tid == 0) { a.trace_E[sci] = E; a.trace_acc[sci] = (int)(nacc - nacc0); }
    }
    for (int j = tid; j < N; j += T_) { P[j] = s.x[j]; P[Npad + j] = s.y[j]; P[2 * Npad + j] = s.z[j]; }
    if (tid == 0) {
        d.E[chain] = E;
        d.nacc[chain] += nacc;
        d.ntri[chain] += (long long)a.nsweeps * N;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
            atomicAdd(d.pair_counts + 1, cnt);
            atomicAdd(d.pair_counts + 2, nscr);
        }
    }
}

template <bool FED>
__global__ void __launch_bounds__(512) k_sweep_block_strict_spec(DevChains d, SweepArgs a)      // 16 warps: the strict arithmetic wants 128 registers
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_block_strict_spec_body<FED, true>(d, a);
    else sweep_block_strict_spec_body<FED, false>(d, a);
}

}  // namespace smcb
