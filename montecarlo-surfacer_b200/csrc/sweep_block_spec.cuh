// sweep_block_spec.cuh — the FAST sweep (oneParticleMoves, SMC.c:278-351) for N > 512, batch-speculative:
// one thread block per chain, one WARP per trial.
//
// k_sweep_block spends a trial on two block-wide evaluations with a barrier and a block reduction each: ~13 600
// cycles per trial at N = 4096 whatever the thread count (profiles/r02).  Here the NW warps of the block
// (32 with 1024 threads) evaluate the next NW trials of the sweep AT THE SAME TIME, each warp on its own: energy and
// force of its molecule at the old position (SMC.c:300-304), the proposal (:307-316), energy and force there
// (:319-321) and the acceptance test (:326-335), all against the positions at the start of the batch - the same
// packed-FP32 superset screen and exact FP64 pair terms as every FAST kernel, summed in a fixed order (ascending
// partner index within a lane, then the warp butterfly).  A trial's result is what the sequential sweep computes
// unless an EARLIER trial of the batch was accepted and moved its molecule out of, into or within the range of the
// trial's old or proposed position.  After a barrier every warp tests exactly that against the earlier trials'
// molecules and proposals (with the inflated screen radius, so the test errs on the safe side); the longest prefix of
// the batch whose results stand is committed in visiting order, and the next batch starts at the first trial that has
// to be redone (trial 0 of a batch has no predecessor: every batch commits at least one trial).  Three barriers per
// BATCH instead of ~5 per trial; rejected trials never void anything, so dense states with low acceptance keep whole
// batches, and a dilute gas rarely has two trials of a batch within range of each other.
#pragma once

namespace smcb {

constexpr int kBlockSpecWords = 6;    // hit bits: 32 molecules per lane and word -> N <= 6144 (shared memory stops at 6016)

struct BlockSpecSmem {
    double *x, *y, *z;        // exact positions                                   [3][Npad]
    double *pq;               // the batch's proposals, exact                      [3][32]
    double *pdU;              // Un - Um of the batch's trials                      [32]
    float *fx, *fy, *fz;      // positions in box units, screen precision          [3][NF], padded with far-away molecules
    float *pf;                // the batch's proposals in box units                [3][32]
    unsigned *px;             // earlier trials of the batch whose acceptance voids this one [32]
    unsigned *pin;            // partners inside the cutoff, old + proposed position [32]
    unsigned *pacc;           // the speculative decisions                         [32]
    static __host__ __device__ int nf(int Npad) { return (Npad + 63) & ~63; }
    __device__ __forceinline__ void carve(double *base, int Npad)
    {
        const int NF = nf(Npad);
        x = base; y = x + Npad; z = y + Npad;
        pq = z + Npad; pdU = pq + 96;
        fx = reinterpret_cast<float *>(pdU + 32); fy = fx + NF; fz = fy + NF;
        pf = fz + NF;
        px = reinterpret_cast<unsigned *>(pf + 96); pin = px + 32; pacc = pin + 32;
    }
    static __host__ __device__ size_t bytes(int Npad)
    {
        return (size_t)(3 * Npad + 128) * sizeof(double) + (size_t)(3 * nf(Npad) + 96) * sizeof(float) + 96 * sizeof(unsigned);
    }
};

// the scalar form of the screen: are two points (box units) within the inflated cutoff?
template <bool PZ>
__device__ __forceinline__ bool screen_near(const ScreenConsts &sc, float ax, float ay, float az, float bx, float by, float bz)
{
    const float MGs = 12582912.f;
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    dx -= (dx + MGs) - MGs;
    dy -= (dy + MGs) - MGs;
    if (PZ) dz = fmaf((dz * sc.inv_zper + MGs) - MGs, -sc.zper, dz);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < sc.rc2s;
}

// one packed screen iteration: a molecule pair (two molecules per float2) against the point; two hit bits are shifted
// into w from the right.  r2 >= 0, so as integers the two floats compare like the numbers do: (bits(r2) - bits(rc2s))
// is negative exactly for a hit, and a funnel shift moves that sign bit into w - two integer instructions per molecule.
template <bool PZ, bool WX, bool WY>
__device__ __forceinline__ void block_spec_screen_pair(const ScreenConsts &sc, float2 mx, float2 my, float2 mz,
                                                       float2 ax, float2 ay, float2 az, int rcbits, unsigned &w)
{
    const float2 MG = make_float2(12582912.f, 12582912.f);
    float2 sx = sub2(ax, mx);
    if (WX) sx = sub2(sx, sub2(add2(sx, MG), MG));
    float2 sy = sub2(ay, my);
    if (WY) sy = sub2(sy, sub2(add2(sy, MG), MG));
    float2 sz = sub2(az, mz);
    if (PZ) {
        const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
        sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
    }
    const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
    w = __funnelshift_l((unsigned)(__float_as_int(r2.x) - rcbits), w, 1);
    w = __funnelshift_l((unsigned)(__float_as_int(r2.y) - rcbits), w, 1);
}

// one pass of the packed-FP32 screen over the chain for NP points q (box units) at once: every molecule pair is loaded
// once and tested against all of them.  Lane l tests the molecule pairs j2 = l + 32 it; 16 iterations fill one 32-bit
// word, FIRST tested molecule in the TOP bit; hq[p][c] holds words 2c (upper half) and 2c + 1 of point p, so that a
// count-leading-zeros walk visits the lane's hits in ascending molecule index.
template <bool PZ, bool WX, bool WY, int NP>
__device__ __forceinline__ void block_spec_screen(const ScreenConsts &sc, const BlockSpecSmem &s, int nit, int lane,
                                                  const float (&qx)[NP], const float (&qy)[NP], const float (&qz)[NP],
                                                  unsigned long long (&hq)[NP][kBlockSpecWords / 2])
{
    float2 ax[NP], ay[NP], az[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        ax[p] = make_float2(qx[p], qx[p]); ay[p] = make_float2(qy[p], qy[p]); az[p] = make_float2(qz[p], qz[p]);
        hq[p][0] = 0ull; hq[p][1] = 0ull; hq[p][2] = 0ull;
    }
    const float2 *X2 = reinterpret_cast<const float2 *>(s.fx), *Y2 = reinterpret_cast<const float2 *>(s.fy),
                 *Z2 = reinterpret_cast<const float2 *>(s.fz);
    const int rcbits = __float_as_int(sc.rc2s);
#pragma unroll 1
    for (int cp = 0; 32 * cp < nit; cp++) {           // two words per pass: the unrolled body stays small (instruction cache)
        unsigned long long v[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) v[p] = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = 2 * cp + h;
            unsigned w[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) w[p] = 0u;
            const int i1 = min(16, nit - 16 * c);
            if (i1 == 16) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int j2 = lane + 32 * (16 * c + i);
                    const float2 mx = X2[j2], my = Y2[j2], mz = Z2[j2];
#pragma unroll
                    for (int p = 0; p < NP; p++) block_spec_screen_pair<PZ, WX, WY>(sc, mx, my, mz, ax[p], ay[p], az[p], rcbits, w[p]);
                }
            } else if (i1 > 0) {                       // the last, partial word of a chain that is no multiple of 1024
                for (int i = 0; i < i1; i++) {
                    const int j2 = lane + 32 * (16 * c + i);
                    const float2 mx = X2[j2], my = Y2[j2], mz = Z2[j2];
#pragma unroll
                    for (int p = 0; p < NP; p++) block_spec_screen_pair<PZ, WX, WY>(sc, mx, my, mz, ax[p], ay[p], az[p], rcbits, w[p]);
                }
#pragma unroll
                for (int p = 0; p < NP; p++) w[p] <<= 32 - 2 * i1;
            }
#pragma unroll
            for (int p = 0; p < NP; p++) v[p] = (v[p] << 32) | (unsigned long long)w[p];
        }
#pragma unroll
        for (int p = 0; p < NP; p++) { if (cp == 0) hq[p][0] = v[p]; else if (cp == 1) hq[p][1] = v[p]; else hq[p][2] = v[p]; }
    }
}

// the next hit of a lane's bit words (ascending molecule index); false when none is left
__device__ __forceinline__ bool block_spec_next_hit(unsigned long long (&hq)[kBlockSpecWords / 2], int lane, int &j)
{
    int o;                                             // order index of the hit: 64 per word, 2 per screen iteration
    if (hq[0]) { const int p = __clzll((long long)hq[0]); hq[0] &= ~(0x8000000000000000ull >> p); o = p; }
    else if (hq[1]) { const int p = __clzll((long long)hq[1]); hq[1] &= ~(0x8000000000000000ull >> p); o = 64 + p; }
    else if (hq[2]) { const int p = __clzll((long long)hq[2]); hq[2] &= ~(0x8000000000000000ull >> p); o = 128 + p; }
    else return false;
    j = 2 * (lane + 32 * (o >> 1)) + (o & 1);
    return true;
}

// energy (already *4) and force of NP molecules `self[p]` placed at p[p], each against all the others and the surface,
// by ONE warp in one pass over the chain; every lane returns the warp totals.  nin[p]: partners inside the cutoff.
template <bool PZ, int NP>
__device__ __forceinline__ void warp_eval_points_smem(const Box &b, const ScreenConsts &sc, const BlockSpecSmem &s, const double *__restrict__ W,
                                                      int N, int nit, int lane, const int (&self)[NP],
                                                      const double (&px)[NP], const double (&py)[NP], const double (&pz)[NP],
                                                      double (&U)[NP], double (&Fx)[NP], double (&Fy)[NP], double (&Fz)[NP], unsigned (&nin)[NP])
{
    float qx[NP], qy[NP], qz[NP];
    bool wx = false, wy = false;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        qx[p] = (float)(px[p] * b.invL); qy[p] = (float)(py[p] * b.invL); qz[p] = (float)(pz[p] * b.invL);
        wx |= !(fabsf(qx[p]) < sc.interior); wy |= !(fabsf(qy[p]) < sc.interior);
    }
    // phase 1: the screen.  The points are the same for the whole warp: the wrap of an axis along which they are
    // interior (ScreenConsts) is dropped from the pass by a uniform branch - the screen is this kernel's load.
    unsigned long long hq[NP][kBlockSpecWords / 2];
    if (wx) {
        if (wy) block_spec_screen<PZ, true, true, NP>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, true, false, NP>(sc, s, nit, lane, qx, qy, qz, hq);
    } else {
        if (wy) block_spec_screen<PZ, false, true, NP>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, false, false, NP>(sc, s, nit, lane, qx, qy, qz, hq);
    }
    // phase 2: the lane's hits in ascending molecule index, exact FP64 terms from the unscaled positions; the
    // points take turns, so that their dependent FP64 chains overlap
    double v[NP][4];
    unsigned mine[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { v[p][0] = 0.0; v[p][1] = 0.0; v[p][2] = 0.0; v[p][3] = 0.0; mine[p] = 0; }
    if (NP == 1) {
        int j;
        while (block_spec_next_hit(hq[0], lane, j)) {
            double et, gx, gy, gz;
            if (j != self[0] && j < N && pair_exact(b, px[0], py[0], pz[0], s.x[j], s.y[j], s.z[j], et, gx, gy, gz)) {
                v[0][0] += et; v[0][1] += gx; v[0][2] += gy; v[0][3] += gz;
                mine[0]++;
            }
        }
    } else for (;;) {
        int j[NP];
        bool has[NP], any = false;
#pragma unroll
        for (int p = 0; p < NP; p++) { has[p] = block_spec_next_hit(hq[p], lane, j[p]); has[p] = has[p] && j[p] != self[p] && j[p] < N; any |= has[p] || hq[p][0] || hq[p][1] || hq[p][2]; }
        if (!any) break;
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const int jj = has[p] ? j[p] : 0;
            double dx, dy, dz;
            const double r2 = pair_sep<false>(b, px[p], py[p], pz[p], s.x[jj], s.y[jj], s.z[jj], dx, dy, dz);
            const bool in = has[p] && r2 < b.rc2;
            const double i2 = fast_rcp(in ? r2 : 1.0);         // same arithmetic as pair_exact
            const double i6 = i2 * i2 * i2;
            const double et = fma(i6, i6, -i6);
            const double g = i2 * i6 * fma(48.0, i6, -24.0);
            if (in) { v[p][0] += et; v[p][1] += g * dx; v[p][2] += g * dy; v[p][3] += g * dz; mine[p]++; }
        }
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < NP; p++) {
        double ew = 0.0, fzw = 0.0;
        if (b.wall) {
            const double dzw = wall_dz<false>(b, pz[p]);
            add_zwall(b, dzw, ew, fzw);                // flat wall: uniform, added after the reduction
            if (dzw * dzw < b.rc2) {                   // surface sites: lane m owns sites m, m + 32, ...
                const int MM = b.M * b.M;
                const double dw = b.L / b.M;
                for (int m = lane; m < MM; m += 32) {
                    const int i = m / b.M, jm = m - i * b.M;
                    const double dx = min_image<false>(px[p] - i * dw, b.L, b.invL);
                    const double dy = min_image<false>(py[p] - jm * dw, b.L, b.invL);
                    const double r2w = fma(dzw, dzw, fma(dy, dy, dx * dx));
                    if (r2w < b.rc2) {
                        const double i2 = fast_rcp(r2w);
                        const double i6 = i2 * i2 * i2;
                        const double a6 = W[2 * m] * i6;
                        v[p][0] += fma(a6, i6, -W[2 * m + 1] * i6);
                        const double g = i2 * i6 * fma(48.0, a6, -24.0 * W[2 * m + 1]);
                        v[p][1] = fma(g, dx, v[p][1]); v[p][2] = fma(g, dy, v[p][2]); v[p][3] = fma(g, dzw, v[p][3]);
                    }
                }
            }
        }
        warp_sum4(lane, v[p][0], v[p][1], v[p][2], v[p][3]);
        nin[p] += __reduce_add_sync(FULL, mine[p]);
        U[p] = 4.0 * (v[p][0] + ew); Fx[p] = v[p][1]; Fy[p] = v[p][2]; Fz[p] = v[p][3] + fzw;
    }
}

// TPW = trials per warp: a batch is NW x TPW <= 32 trials, trial t of the batch belongs to warp t % NW, slot t / NW
template <bool FED, bool PZ, int TPW>
__device__ __forceinline__ void sweep_block_spec_body(const DevChains &d, const SweepArgs &a)
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

    const double AoT = b.A / b.T, sigma = sqrt(2.0 * b.A), quarterAoT = 0.25 * AoT, invT = 1.0 / b.T;
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    double E = d.E[chain];                             // kept by thread 0
    int nacc = 0;
    unsigned long long cnt = 0, nscr = 0;              // thread 0
    const int BMAX = min(32, NW * TPW);

    for (int sw = 0; sw < a.nsweeps; sw++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)sw;
        const size_t sci = (size_t)sw * d.C + chain;
        const int nacc0 = nacc;
        long long offset;                              // int offset = rand();  SMC.c:290
        if (FED) offset = a.offset[sci];
        else { uint32_t o; double unused; rng_step_scalars(id, step, o, unused); offset = o; }
        const int off = (int)(offset % N);
        int nn0 = 0;
        while (nn0 < N) {
            const int Bn = min(BMAX, N - nn0);         // trials nn0 .. nn0 + Bn - 1 of the sweep
            int n[TPW], tix[TPW];
            bool act[TPW], acc[TPW];
            double px[TPW], py[TPW], pz[TPW], qx[TPW], qy[TPW], qz[TPW];
            float ox[TPW], oy[TPW], oz[TPW], nx[TPW], ny[TPW], nz[TPW];
#pragma unroll
            for (int p = 0; p < TPW; p++) {
                tix[p] = warp + NW * p;
                act[p] = tix[p] < Bn;
                n[p] = nn0 + (act[p] ? tix[p] : 0) + off;      // n = (nn+offset)%N  SMC.c:294; idle slots shadow trial 0
                if (n[p] >= N) n[p] -= N;
                px[p] = s.x[n[p]]; py[p] = s.y[n[p]]; pz[p] = s.z[n[p]];
                acc[p] = false;
            }
            if (act[0]) {
                double g0[TPW], g1[TPW], g2[TPW], lul[TPW];    // every lane of the warp forms the same numbers
#pragma unroll
                for (int p = 0; p < TPW; p++) {
                    double ul;
                    if (FED) {
                        const double *dsp = a.displ + sci * 3 * N;
                        g0[p] = dsp[3 * n[p]]; g1[p] = dsp[3 * n[p] + 1]; g2[p] = dsp[3 * n[p] + 2];
                        ul = a.u[sci * N + nn0 + (act[p] ? tix[p] : 0)];
                    } else {
                        rng_particle_gauss_f32(id, step, (uint32_t)n[p], g0[p], g1[p], g2[p]);
                        g0[p] *= sigma; g1[p] *= sigma; g2[p] *= sigma;
                        ul = rng_particle_uniform(id, step, (uint32_t)n[p]);
                    }
                    lul[p] = log(ul);
                }
                double Um[TPW], Fmx[TPW], Fmy[TPW], Fmz[TPW], Un[TPW], Fnx[TPW], Fny[TPW], Fnz[TPW], dX[TPW], dY[TPW], dZ[TPW];
                unsigned nin[TPW];
#pragma unroll
                for (int p = 0; p < TPW; p++) nin[p] = 0;
                warp_eval_points_smem<PZ, TPW>(b, sc, s, W, N, nit, lane, n, px, py, pz, Um, Fmx, Fmy, Fmz, nin);      // SMC.c:300-304
#pragma unroll
                for (int p = 0; p < TPW; p++) {
                    dX[p] = fma(Fmx[p], AoT, g0[p]); dY[p] = fma(Fmy[p], AoT, g1[p]); dZ[p] = fma(Fmz[p], AoT, g2[p]);   // SMC.c:307-309
                    qx[p] = min_image<false>(px[p] + dX[p], b.L, b.invL); qy[p] = min_image<false>(py[p] + dY[p], b.L, b.invL);   // SMC.c:311-316
                    qz[p] = pz[p] + dZ[p];
                    if (PZ) qz[p] = min_image<false>(qz[p], b.Lz, b.invLz);
                }
                warp_eval_points_smem<PZ, TPW>(b, sc, s, W, N, nit, lane, n, qx, qy, qz, Un, Fnx, Fny, Fnz, nin);      // SMC.c:319-321
#pragma unroll
                for (int p = 0; p < TPW; p++) {
                    // SMC.c:326-335: accept iff u < exp(-(Un-Um + d.(Fn+Fm)/2 + (Fn^2-Fm^2) A/(4T))/T)
                    const double f2 = fma(Fnx[p], Fnx[p], fma(Fny[p], Fny[p], Fnz[p] * Fnz[p])) - fma(Fmx[p], Fmx[p], fma(Fmy[p], Fmy[p], Fmz[p] * Fmz[p]));
                    const double dr = fma(dX[p], Fnx[p] + Fmx[p], fma(dY[p], Fny[p] + Fmy[p], dZ[p] * (Fnz[p] + Fmz[p])));
                    const double xarg = -((Un[p] - Um[p]) + 0.5 * dr + f2 * quarterAoT) * invT;
                    acc[p] = act[p] && (lul[p] < xarg) && (xarg > -745.1332191019411);
                    ox[p] = s.fx[n[p]]; oy[p] = s.fy[n[p]]; oz[p] = s.fz[n[p]];
                    nx[p] = (float)(qx[p] * b.invL); ny[p] = (float)(qy[p] * b.invL); nz[p] = (float)(qz[p] * b.invL);
                    // the molecules of the EARLIER trials of the batch, where they are now: in range of my old or proposed position?
                    bool hit = false;
                    if (lane < tix[p]) {
                        int nm = nn0 + lane + off;
                        if (nm >= N) nm -= N;
                        const float mx = s.fx[nm], my = s.fy[nm], mz = s.fz[nm];
                        hit = screen_near<PZ>(sc, ox[p], oy[p], oz[p], mx, my, mz) || screen_near<PZ>(sc, nx[p], ny[p], nz[p], mx, my, mz);
                    }
                    const unsigned x1 = __ballot_sync(FULL, hit);
                    if (lane == 0 && act[p]) {
                        const int t = tix[p];
                        s.pq[t] = qx[p]; s.pq[32 + t] = qy[p]; s.pq[64 + t] = qz[p];
                        s.pf[t] = nx[p]; s.pf[32 + t] = ny[p]; s.pf[64 + t] = nz[p];
                        s.pdU[t] = Un[p] - Um[p];
                        s.pacc[t] = acc[p] ? 1u : 0u; s.pin[t] = nin[p]; s.px[t] = x1;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int p = 0; p < TPW; p++) {
                if (act[p]) {                          // ... and where they would move to
                    bool hit = false;
                    if (lane < tix[p]) {
                        const float mx = s.pf[lane], my = s.pf[32 + lane], mz = s.pf[64 + lane];
                        hit = screen_near<PZ>(sc, ox[p], oy[p], oz[p], mx, my, mz) || screen_near<PZ>(sc, nx[p], ny[p], nz[p], mx, my, mz);
                    }
                    const unsigned x2 = __ballot_sync(FULL, hit);
                    if (lane == 0) s.px[tix[p]] |= x2;
                }
            }
            __syncthreads();
            // every trial before f saw exactly the state the sequential sweep shows it: f = the first trial with an
            // accepted predecessor in range (its predecessors are all before f, so their decisions are final)
            const unsigned accmask = __ballot_sync(FULL, lane < Bn && s.pacc[lane] != 0u);
            const unsigned bad = __ballot_sync(FULL, lane < Bn && (s.px[lane] & accmask) != 0u);
            const int f = bad ? __ffs(bad) - 1 : Bn;
#pragma unroll
            for (int p = 0; p < TPW; p++) {
                if (tix[p] < f && acc[p] && lane == 0) {
                    s.x[n[p]] = qx[p]; s.y[n[p]] = qy[p]; s.z[n[p]] = qz[p];
                    s.fx[n[p]] = nx[p]; s.fy[n[p]] = ny[p]; s.fz[n[p]] = nz[p];
                }
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
            __syncthreads();                            // the moves are visible before the next batch reads positions
            nn0 += f;
        }
        if (a.trace_E != nullptr && tid == 0) { a.trace_E[sci] = E; a.trace_acc[sci] = nacc - nacc0; }
    }

    for (int j = tid; j < N; j += T_) { P[j] = s.x[j]; P[Npad + j] = s.y[j]; P[2 * Npad + j] = s.z[j]; }
    if (tid == 0) {
        d.E[chain] = E;
        d.nacc[chain] += nacc;
        d.ntri[chain] += (long long)a.nsweeps * N;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
            atomicAdd(d.pair_counts + 1, cnt);
            atomicAdd(d.pair_counts + 2, nscr);         // pair tests executed, redone trials included
        }
    }
}

// TPW = 1 (default): 1024 threads, one trial per warp (64 registers).  TPW = 2 (SMCB_BLOCK_SPEC_TPW=2): 512 threads, two
// trials per warp (128 registers), every molecule pair loaded once for both - half the shared-memory wavefronts, the
// same instruction count, and 30 % slower: 4 warps per scheduler hide less latency than 8 (profiles/r02)
template <bool FED, int TPW>
__global__ void __launch_bounds__(1024 / TPW) k_sweep_block_spec(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_block_spec_body<FED, true, TPW>(d, a);
    else sweep_block_spec_body<FED, false, TPW>(d, a);
}

}  // namespace smcb
