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

// one packed screen iteration: the molecule pair j2 against the point; two hit bits are shifted into w from the right.
// r2 >= 0, so as integers the two floats compare like the numbers do: (bits(r2) - bits(rc2s)) is negative exactly for a
// hit, and a funnel shift moves that sign bit into w - two integer instructions per molecule, no predicates.
template <bool PZ, bool WX, bool WY>
__device__ __forceinline__ void block_spec_screen_pair(const ScreenConsts &sc, const float2 *X2, const float2 *Y2, const float2 *Z2, int j2,
                                                       float2 ax, float2 ay, float2 az, int rcbits, unsigned &w)
{
    const float2 MG = make_float2(12582912.f, 12582912.f);
    float2 sx = sub2(ax, X2[j2]);
    if (WX) sx = sub2(sx, sub2(add2(sx, MG), MG));
    float2 sy = sub2(ay, Y2[j2]);
    if (WY) sy = sub2(sy, sub2(add2(sy, MG), MG));
    float2 sz = sub2(az, Z2[j2]);
    if (PZ) {
        const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
        sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
    }
    const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
    w = __funnelshift_l((unsigned)(__float_as_int(r2.x) - rcbits), w, 1);
    w = __funnelshift_l((unsigned)(__float_as_int(r2.y) - rcbits), w, 1);
}

// one pass of the packed-FP32 screen over the chain for the point q (box units).  Lane l tests the molecule pairs
// j2 = l + 32 it; 16 iterations fill one 32-bit word, FIRST tested molecule in the TOP bit; hq[c] holds words 2c
// (upper half) and 2c + 1, so that a count-leading-zeros walk visits the lane's hits in ascending molecule index.
template <bool PZ, bool WX, bool WY>
__device__ __forceinline__ void block_spec_screen(const ScreenConsts &sc, const BlockSpecSmem &s, int nit, int lane,
                                                  float qx, float qy, float qz, unsigned long long (&hq)[kBlockSpecWords / 2])
{
    const float2 ax = make_float2(qx, qx), ay = make_float2(qy, qy), az = make_float2(qz, qz);
    const float2 *X2 = reinterpret_cast<const float2 *>(s.fx), *Y2 = reinterpret_cast<const float2 *>(s.fy),
                 *Z2 = reinterpret_cast<const float2 *>(s.fz);
    const int rcbits = __float_as_int(sc.rc2s);
    hq[0] = 0ull; hq[1] = 0ull; hq[2] = 0ull;
#pragma unroll 1
    for (int cp = 0; 32 * cp < nit; cp++) {           // two words per pass: the unrolled body stays small (instruction cache)
        unsigned long long v = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = 2 * cp + h;
            unsigned w = 0;
            const int i1 = min(16, nit - 16 * c);
            if (i1 == 16) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    block_spec_screen_pair<PZ, WX, WY>(sc, X2, Y2, Z2, lane + 32 * (16 * c + i), ax, ay, az, rcbits, w);
            } else if (i1 > 0) {                       // the last, partial word of a chain that is no multiple of 1024
                for (int i = 0; i < i1; i++)
                    block_spec_screen_pair<PZ, WX, WY>(sc, X2, Y2, Z2, lane + 32 * (16 * c + i), ax, ay, az, rcbits, w);
                w <<= 32 - 2 * i1;
            }
            v = (v << 32) | (unsigned long long)w;
        }
        if (cp == 0) hq[0] = v; else if (cp == 1) hq[1] = v; else hq[2] = v;
    }
}

// energy (already *4) and force of molecule `self` placed at p, against all the others and the surface, by ONE warp;
// every lane returns the warp totals.  nin: partners inside the cutoff (warp total).
template <bool PZ>
__device__ __forceinline__ void warp_eval_point_smem(const Box &b, const ScreenConsts &sc, const BlockSpecSmem &s, const double *__restrict__ W,
                                                     int N, int nit, int self, double px, double py, double pz, int lane,
                                                     double &U, double &Fx, double &Fy, double &Fz, unsigned &nin)
{
    const float qx = (float)(px * b.invL), qy = (float)(py * b.invL), qz = (float)(pz * b.invL);
    // phase 1: the screen.  Lane l tests the molecule pairs j2 = l + 32 it (molecules 2 j2, 2 j2 + 1), 16 iterations
    // fill one word of hit bits; nit = NF / 64 iterations in all, the same for every lane (the arrays are padded).
    // The point is the same for the whole warp: the wrap of an axis along which it is interior (ScreenConsts) is
    // dropped from the pass by a uniform branch - the screen is the packed-FP32 pipe's load of this kernel.
    unsigned long long hq[kBlockSpecWords / 2];
    const bool wx = !(fabsf(qx) < sc.interior), wy = !(fabsf(qy) < sc.interior);
    if (wx) {
        if (wy) block_spec_screen<PZ, true, true>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, true, false>(sc, s, nit, lane, qx, qy, qz, hq);
    } else {
        if (wy) block_spec_screen<PZ, false, true>(sc, s, nit, lane, qx, qy, qz, hq);
        else block_spec_screen<PZ, false, false>(sc, s, nit, lane, qx, qy, qz, hq);
    }
    // phase 2: the lane's hits in ascending molecule index, exact FP64 terms from the unscaled positions
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    unsigned mine = 0;
    for (;;) {
        int o;                                         // order index of the hit: 64 per hq word, 2 per screen iteration
        if (hq[0]) { const int p = __clzll((long long)hq[0]); hq[0] &= ~(0x8000000000000000ull >> p); o = p; }
        else if (hq[1]) { const int p = __clzll((long long)hq[1]); hq[1] &= ~(0x8000000000000000ull >> p); o = 64 + p; }
        else if (hq[2]) { const int p = __clzll((long long)hq[2]); hq[2] &= ~(0x8000000000000000ull >> p); o = 128 + p; }
        else break;
        const int j = 2 * (lane + 32 * (o >> 1)) + (o & 1);
        double et, gx, gy, gz;
        if (j != self && j < N && pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, gx, gy, gz)) {
            v0 += et; v1 += gx; v2 += gy; v3 += gz;
            mine++;
        }
    }
    __syncwarp();
    double ew = 0.0, fzw = 0.0;
    if (b.wall) {
        const double dzw = wall_dz<false>(b, pz);
        add_zwall(b, dzw, ew, fzw);                    // flat wall: uniform, added after the reduction
        if (dzw * dzw < b.rc2) {                       // surface sites: lane m owns sites m, m + 32, ...
            const int MM = b.M * b.M;
            const double dw = b.L / b.M;
            for (int m = lane; m < MM; m += 32) {
                const int i = m / b.M, j = m - i * b.M;
                const double dx = min_image<false>(px - i * dw, b.L, b.invL);
                const double dy = min_image<false>(py - j * dw, b.L, b.invL);
                const double r2w = fma(dzw, dzw, fma(dy, dy, dx * dx));
                if (r2w < b.rc2) {
                    const double i2 = fast_rcp(r2w);
                    const double i6 = i2 * i2 * i2;
                    const double a6 = W[2 * m] * i6;
                    v0 += fma(a6, i6, -W[2 * m + 1] * i6);
                    const double g = i2 * i6 * fma(48.0, a6, -24.0 * W[2 * m + 1]);
                    v1 = fma(g, dx, v1); v2 = fma(g, dy, v2); v3 = fma(g, dzw, v3);
                }
            }
        }
    }
    warp_sum4(lane, v0, v1, v2, v3);
    nin += __reduce_add_sync(FULL, mine);
    U = 4.0 * (v0 + ew); Fx = v1; Fy = v2; Fz = v3 + fzw;
}

template <bool FED, bool PZ>
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
            const int Bn = min(NW, N - nn0);           // trials nn0 .. nn0 + Bn - 1, warp w takes trial nn0 + w
            bool acc = false;
            double qx = 0.0, qy = 0.0, qz = 0.0;
            float ox = 0.f, oy = 0.f, oz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
            int n = 0;
            if (warp < Bn) {
                const int nn = nn0 + warp;
                n = nn + off;                          // n = (nn+offset)%N  SMC.c:294
                if (n >= N) n -= N;
                double g0, g1, g2, ul;                 // every lane of the warp forms the same numbers
                if (FED) {
                    const double *dsp = a.displ + sci * 3 * N;
                    g0 = dsp[3 * n]; g1 = dsp[3 * n + 1]; g2 = dsp[3 * n + 2];
                    ul = a.u[sci * N + nn];
                } else {
                    rng_particle_gauss_f32(id, step, (uint32_t)n, g0, g1, g2);
                    g0 *= sigma; g1 *= sigma; g2 *= sigma;
                    ul = rng_particle_uniform(id, step, (uint32_t)n);
                }
                const double px = s.x[n], py = s.y[n], pz = s.z[n];
                double Um, Fmx, Fmy, Fmz, Un, Fnx, Fny, Fnz;
                unsigned nin = 0;
                warp_eval_point_smem<PZ>(b, sc, s, W, N, nit, n, px, py, pz, lane, Um, Fmx, Fmy, Fmz, nin);          // SMC.c:300-304
                const double dX = fma(Fmx, AoT, g0), dY = fma(Fmy, AoT, g1), dZ = fma(Fmz, AoT, g2);               // SMC.c:307-309
                qx = min_image<false>(px + dX, b.L, b.invL); qy = min_image<false>(py + dY, b.L, b.invL);           // SMC.c:311-316
                qz = pz + dZ;
                if (PZ) qz = min_image<false>(qz, b.Lz, b.invLz);
                warp_eval_point_smem<PZ>(b, sc, s, W, N, nit, n, qx, qy, qz, lane, Un, Fnx, Fny, Fnz, nin);          // SMC.c:319-321
                // SMC.c:326-335: accept iff u < exp(-(Un-Um + d.(Fn+Fm)/2 + (Fn^2-Fm^2) A/(4T))/T)
                const double f2 = fma(Fnx, Fnx, fma(Fny, Fny, Fnz * Fnz)) - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                const double dr = fma(dX, Fnx + Fmx, fma(dY, Fny + Fmy, dZ * (Fnz + Fmz)));
                const double xarg = -((Un - Um) + 0.5 * dr + f2 * quarterAoT) * invT;
                acc = (log(ul) < xarg) && (xarg > -745.1332191019411);
                ox = s.fx[n]; oy = s.fy[n]; oz = s.fz[n];
                nx = (float)(qx * b.invL); ny = (float)(qy * b.invL); nz = (float)(qz * b.invL);
                // the molecules of the EARLIER trials of the batch, where they are now: in range of my old or proposed position?
                bool hit = false;
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
                    s.pacc[warp] = acc ? 1u : 0u; s.pin[warp] = nin; s.px[warp] = x1;
                }
            }
            __syncthreads();
            if (warp < Bn) {                           // ... and where they would move to
                bool hit = false;
                if (lane < warp) {
                    const float mx = s.pf[lane], my = s.pf[32 + lane], mz = s.pf[64 + lane];
                    hit = screen_near<PZ>(sc, ox, oy, oz, mx, my, mz) || screen_near<PZ>(sc, nx, ny, nz, mx, my, mz);
                }
                const unsigned x2 = __ballot_sync(FULL, hit);
                if (lane == 0) s.px[warp] |= x2;
            }
            __syncthreads();
            // every trial before f saw exactly the state the sequential sweep shows it: f = the first trial with an
            // accepted predecessor in range (its predecessors are all before f, so their decisions are final)
            const unsigned accmask = __ballot_sync(FULL, lane < Bn && s.pacc[lane] != 0u);
            const unsigned bad = __ballot_sync(FULL, lane < Bn && (s.px[lane] & accmask) != 0u);
            const int f = bad ? __ffs(bad) - 1 : Bn;
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

#ifndef SMCB_BLOCK_SPEC_THREADS
#define SMCB_BLOCK_SPEC_THREADS 1024
#endif
template <bool FED>
__global__ void __launch_bounds__(SMCB_BLOCK_SPEC_THREADS) k_sweep_block_spec(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_block_spec_body<FED, true>(d, a);
    else sweep_block_spec_body<FED, false>(d, a);
}

}  // namespace smcb
