// sweep_cached.cuh — the FAST sweep kernel (oneParticleMoves, SMC.c:278-351).
//
// Same Markov chain as the reference's sweep (same proposal, same acceptance
// expression, same visiting order, same random inputs), organised for the GPU:
//
//  * one warp per chain; lane l keeps particles j = l + 32k in registers, in box
//    units (x/L), so a minimum image is s - rint(s) and a pair costs 9 FP64-pipe
//    instructions (phase 1: geometry + cutoff screen only);
//  * per-particle energy e_i (= energySingle + wallsEnergySingle), force F_i
//    (= forceSingle + wallsForce) and LJ neighbour count are CACHED in shared memory.
//    The reference recomputes the old-position energy/force of the trial particle
//    from scratch (SMC.c:300-304); here they are read from the cache, so a trial needs
//    ONE O(N) pass (the proposed position) and one warp reduction instead of two.
//    When a move is accepted the caches of the partners inside the old/new cutoff
//    spheres are corrected by the pair terms (Newton's third law); the old-position
//    pass that finds the old partners is skipped when the cached neighbour count is 0
//    (most trials in the gas phase);
//  * pairs inside the cutoff ("hits") are rare in the gas and are handled in a
//    divergent phase 2 from the exact (unscaled) shared-memory mirror, with the same
//    arithmetic as k_evaluate's FAST path; the in/out decision is a symmetric,
//    deterministic function of the two positions, so a pair is always removed from
//    the caches by the same test that added it (the neighbour counts stay exact).
//
// The caches are rebuilt from the positions at the start of every launch, which
// bounds the rounding drift of the incremental updates (a launch is tens of sweeps).
// Results agree with the reference to ~1e-15 per trial (tests: teacher-forced 1e-12);
// the bit-exact path is k_sweep<.., STRICT>.
#pragma once

namespace smcb {

// shared-memory carve-up of one chain
struct ChainSmem {
    double *x, *y, *z;        // exact positions (mirror of the registers)
    double *ce, *cfx, *cfy, *cfz;   // cached per-particle energy and force (LJ + surface)
    unsigned short *nb;       // cached number of LJ partners inside the cutoff
    double *site;             // [4][MMpad]: site x, y, a, b
    __device__ __forceinline__ void carve(double *base, int Npad, int MMpad)
    {
        x = base; y = x + Npad; z = y + Npad;
        ce = z + Npad; cfx = ce + Npad; cfy = cfx + Npad; cfz = cfy + Npad;
        site = cfz + Npad;
        nb = reinterpret_cast<unsigned short *>(site + 4 * MMpad);
    }
    static __host__ __device__ size_t bytes(int Npad, int MMpad)
    {
        return (size_t)(7 * Npad + 4 * MMpad) * sizeof(double) + (size_t)Npad * sizeof(unsigned short);
    }
};

struct ScreenConsts {
    double rc2s;              // rc2 / L^2, inflated by 1e-12 (phase-1 screen; exact test in phase 2)
    double zper, inv_zper;    // Lz/L and L/Lz (bulk mode only)
};

// phase 1: which of the lane's K slots are within the (slightly inflated) cutoff of the
// point (psx,psy,psz) given in box units
template <int K, bool PZ>
__device__ __forceinline__ unsigned screen_slots(const Box &b, const ScreenConsts &sc, double psx, double psy, double psz,
                                                 const double (&xs)[K], const double (&ys)[K], const double (&zs)[K])
{
    unsigned hits = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double sx = wrap_unit_x(psx - xs[k]);
        const double sy = wrap_unit_y(psy - ys[k]);
        double sz = psz - zs[k];
        if (PZ) sz = fma(-sc.zper, rint(sz * sc.inv_zper), sz);
        const double r2s = fma(sz, sz, fma(sy, sy, sx * sx));
        if (r2s < sc.rc2s) hits |= 1u << k;
    }
    return hits;
}

// two points against the same slots in one loop (old and proposed position): 16
// independent pair evaluations for the scheduler to interleave
template <int K, bool PZ>
__device__ __forceinline__ void screen_slots2(const Box &b, const ScreenConsts &sc,
                                              double ax, double ay, double az, double bx, double by, double bz,
                                              const double (&xs)[K], const double (&ys)[K], const double (&zs)[K],
                                              unsigned &hits_a, unsigned &hits_b)
{
    hits_a = hits_b = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double sxa = wrap_unit_x(ax - xs[k]), sxb = wrap_unit_x(bx - xs[k]);
        const double sya = wrap_unit_y(ay - ys[k]), syb = wrap_unit_y(by - ys[k]);
        double sza = az - zs[k], szb = bz - zs[k];
        if (PZ) {
            sza = fma(-sc.zper, rint(sza * sc.inv_zper), sza);
            szb = fma(-sc.zper, rint(szb * sc.inv_zper), szb);
        }
        if (fma(sza, sza, fma(sya, sya, sxa * sxa)) < sc.rc2s) hits_a |= 1u << k;
        if (fma(szb, szb, fma(syb, syb, sxb * sxb)) < sc.rc2s) hits_b |= 1u << k;
    }
}

// exact 12-6 terms of one pair from unscaled positions (same arithmetic as lj_terms<false,true>
// with the call-free reciprocal); returns false when the pair is outside the true cutoff
__device__ __forceinline__ bool pair_exact(const Box &b, double px, double py, double pz, double jx, double jy, double jz,
                                           double &e, double &gx, double &gy, double &gz)
{
    double dx, dy, dz;
    const double r2 = pair_sep<false>(b, px, py, pz, jx, jy, jz, dx, dy, dz);
    if (!(r2 < b.rc2)) return false;
    const double i2 = fast_rcp(r2);
    const double i6 = i2 * i2 * i2;
    e = fma(i6, i6, -i6);
    const double g = i2 * i6 * fma(48.0, i6, -24.0);
    gx = g * dx; gy = g * dy; gz = g * dz;
    return true;
}

// phase 2 for the point p: add the terms of the lane's hits; returns the mask of slots
// that are truly inside the cutoff
__device__ __forceinline__ unsigned add_hits(const Box &b, const ChainSmem &s, int lane, unsigned hits,
                                             double px, double py, double pz,
                                             double &e, double &fx, double &fy, double &fz)
{
    unsigned in = 0;
    while (hits) {
        const int k = __ffs(hits) - 1;
        hits &= hits - 1;
        const int j = lane + 32 * k;
        double et, gx, gy, gz;
        if (pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, gx, gy, gz)) {
            e += et; fx += gx; fy += gy; fz += gz;
            in |= 1u << k;
        }
    }
    return in;
}

// surface sites for the point p (lane m < M*M owns site m; more sites loop), flat wall excluded
__device__ __forceinline__ void add_sites(const Box &b, const ChainSmem &s, int lane, int MMpad,
                                          double px, double py, double dzw,
                                          double &e, double &fx, double &fy, double &fz)
{
    const int MM = b.M * b.M;
    for (int m = lane; m < MM; m += 32) {
        const double dx = min_image<false>(px - s.site[m], b.L, b.invL);
        const double dy = min_image<false>(py - s.site[MMpad + m], b.L, b.invL);
        const double r2 = fma(dzw, dzw, fma(dy, dy, dx * dx));
        if (r2 < b.rc2) {
            const double ca = s.site[2 * MMpad + m], cb = s.site[3 * MMpad + m];
            const double i2 = fast_rcp(r2);
            const double i6 = i2 * i2 * i2;
            const double a6 = ca * i6;
            e += fma(a6, i6, -cb * i6);
            const double g = i2 * i6 * fma(48.0, a6, -24.0 * cb);
            fx = fma(g, dx, fx);
            fy = fma(g, dy, fy);
            fz = fma(g, dzw, fz);
        }
    }
}

// flat wall a0/dz^12 - b0/dz^6 (no cutoff, SMC.c:740-741, 787-789); uniform across the warp
__device__ __forceinline__ void add_zwall(const Box &b, double dzw, double &e, double &fz)
{
    const double i2 = fast_rcp(dzw * dzw);
    const double i6 = i2 * i2 * i2;
    const double a6 = b.a0 * i6;
    e += fma(a6, i6, -b.b0 * i6);
    fz = fma(i2 * i6 * fma(48.0, a6, -24.0 * b.b0), dzw, fz);
}

// energy (already *4) and force of a particle at p against everything else; `in`
// receives the lane's exact in-cutoff slots.  All lanes return the warp totals.
template <int K, bool PZ>
__device__ __forceinline__ void eval_point(const Box &b, const ScreenConsts &sc, const ChainSmem &s, int lane, int MMpad,
                                           unsigned okmask, double px, double py, double pz,
                                           const double (&xs)[K], const double (&ys)[K], const double (&zs)[K],
                                           double &U, double &Fx, double &Fy, double &Fz, unsigned &in)
{
    unsigned hits = screen_slots<K, PZ>(b, sc, px * b.invL, py * b.invL, pz * b.invL, xs, ys, zs) & okmask;
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    in = add_hits(b, s, lane, hits, px, py, pz, e, fx, fy, fz);
    double dzw = 0.0;
    if (b.wall) {
        dzw = wall_dz<false>(b, pz);
        if (dzw * dzw < b.rc2) add_sites(b, s, lane, MMpad, px, py, dzw, e, fx, fy, fz);
    }
    warp_sum4(lane, e, fx, fy, fz);
    if (b.wall) add_zwall(b, dzw, e, fz);
    U = 4.0 * e; Fx = fx; Fy = fy; Fz = fz;
}

#ifndef SMCB_SWEEP_MINB
#define SMCB_SWEEP_MINB 10
#endif
template <int K, bool FED, bool PZ>
__device__ __forceinline__ void sweep_cached_body(const DevChains &d, const SweepArgs &a)
{
    const int lane = threadIdx.x, chain = blockIdx.x;
    const int N = d.N, Npad = d.Npad;
    const int MM = d.M * d.M, MMpad = (MM + 3) & ~3;
    extern __shared__ double sm[];
    ChainSmem s;
    s.carve(sm, Npad, MMpad);
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * MM;
    double *P = d.pos + (size_t)chain * 3 * Npad;

    double xs[K], ys[K], zs[K];
    unsigned validmask = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        const bool in = j < N;
        const double X = in ? P[j] : 0.0, Y = in ? P[Npad + j] : 0.0, Z = in ? P[2 * Npad + j] : 0.0;
        if (j < Npad) { s.x[j] = X; s.y[j] = Y; s.z[j] = Z; }
        xs[k] = X * b.invL; ys[k] = Y * b.invL; zs[k] = Z * b.invL;
        if (in) validmask |= 1u << k;
    }
    if (b.wall) {
        const double dw = b.L / d.M;
        for (int m = lane; m < MM; m += 32) {
            const int si = m / d.M, sj = m - si * d.M;
            s.site[m] = si * dw; s.site[MMpad + m] = sj * dw;
            s.site[2 * MMpad + m] = W[2 * m]; s.site[3 * MMpad + m] = W[2 * m + 1];
        }
    }
    __syncwarp();
    ScreenConsts sc;
    sc.rc2s = b.rc2 * b.invL * b.invL * (1.0 + 1e-12);
    sc.zper = b.Lz * b.invL; sc.inv_zper = b.L * b.invLz;

    // ---- rebuild the caches from the positions ------------------------------------
    for (int n = 0; n < N; n++) {
        const unsigned okmask = validmask & ~(((n & 31) == lane) ? (1u << (n >> 5)) : 0u);
        double U, Fx, Fy, Fz;
        unsigned in;
        eval_point<K, PZ>(b, sc, s, lane, MMpad, okmask, s.x[n], s.y[n], s.z[n], xs, ys, zs, U, Fx, Fy, Fz, in);
        const int cntn = __reduce_add_sync(FULL, __popc(in));
        if (lane == 0) { s.ce[n] = U; s.cfx[n] = Fx; s.cfy[n] = Fy; s.cfz[n] = Fz; s.nb[n] = (unsigned short)cntn; }
    }
    __syncwarp();

    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    const double quarterAoT = 0.25 * AoT, invT = 1.0 / b.T;
    double E = d.E[chain];
    int nacc = 0;
    unsigned cnt = 0;                                // per-lane, < 2^32 per launch
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};

    // Physical register slot 0 always holds the slot that is being visited: the visiting
    // order n = (nn+offset)%N walks the slots cyclically, so the register arrays are rotated
    // by one at every slot boundary (48 moves per 32 trials) instead of being updated through
    // a run-time index at every accepted trial.  `rot` = logical slot held at physical 0.
    int rot = 0;
    auto rotate = [&]() {
        if (K > 1) {
            const double tx = xs[0], ty = ys[0], tz = zs[0];
#pragma unroll
            for (int k = 0; k + 1 < K; k++) { xs[k] = xs[k + 1]; ys[k] = ys[k + 1]; zs[k] = zs[k + 1]; }
            xs[K - 1] = tx; ys[K - 1] = ty; zs[K - 1] = tz;
            validmask = (validmask >> 1) | ((validmask & 1u) << (K - 1));
            rot = (rot + 1 == K) ? 0 : rot + 1;
        }
    };
    // physical slot k holds logical slot (k + rot) mod K
    auto particle_of = [&](int k) { int sl = k + rot; if (sl >= K) sl -= K; return lane + 32 * sl; };

    for (int sw = 0; sw < a.nsweeps; sw++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)sw;
        const size_t sci = (size_t)sw * d.C + chain;
        const int nacc0 = nacc;
        long long offset;                              // int offset = rand();  SMC.c:290
        if (FED) {
            offset = a.offset[sci];
        } else {
            uint32_t o; double unused;
            rng_step_scalars(id, step, o, unused);
            offset = o;
        }
        const int off = (int)(offset % N);             // first particle of the sweep: n = (nn+offset)%N, SMC.c:294
        const int slot0 = off >> 5, t0 = off & 31;
        while (rot != slot0) rotate();
        // K+1 segments: [off .. end of its slot], the following slots cyclically, then [start of slot0 .. off-1]
        for (int seg = 0; seg <= K; seg++) {
            const int slot = rot;
            const int tb = (seg == 0) ? t0 : 0;
            int te = min(32, N - 32 * slot);            // particles of this slot that exist
            if (seg == K) te = min(te, t0);
            if (tb < te) {
                // each lane prepares the random inputs of its own particle of this slot
                const int nl = 32 * slot + lane;
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, lul = 0.0;
                if (lane >= tb && lane < te) {
                    double ul;
                    if (FED) {
                        const double *dsp = a.displ + sci * 3 * N;
                        g0 = dsp[3 * nl]; g1 = dsp[3 * nl + 1]; g2 = dsp[3 * nl + 2];
                        int nn = nl - off;              // trial index in visiting order (u is per trial, SMC.c:335)
                        if (nn < 0) nn += N;
                        ul = a.u[sci * N + nn];
                    } else {
                        rng_particle_gauss(id, step, (uint32_t)nl, g0, g1, g2);
                        g0 *= sigma; g1 *= sigma; g2 *= sigma;
                        ul = rng_particle_uniform(id, step, (uint32_t)nl);
                    }
                    lul = log(ul);                      // u < exp(x)  <=>  log(u) < x, evaluated lane-parallel
                }
                for (int t = tb; t < te; t++) {
                    const int n = 32 * slot + t;
                    const unsigned okmask = validmask & ~((lane == t) ? 1u : 0u);
                    const int nbm = s.nb[n];
                    double dX, dY, dZ, qx, qy, qz;
                    {
                        // proposal from the cached force of particle n (SMC.c:303-316)
                        dX = fma(s.cfx[n], AoT, __shfl_sync(FULL, g0, t));
                        dY = fma(s.cfy[n], AoT, __shfl_sync(FULL, g1, t));
                        dZ = fma(s.cfz[n], AoT, __shfl_sync(FULL, g2, t));
                        qx = min_image<false>(s.x[n] + dX, b.L, b.invL);
                        qy = min_image<false>(s.y[n] + dY, b.L, b.invL);
                        qz = s.z[n] + dZ;
                        if (PZ) qz = min_image<false>(qz, b.Lz, b.invLz);
                    }
                    const double qsx = qx * b.invL, qsy = qy * b.invL, qsz = qz * b.invL;

                    // flat wall at the proposal: uniform, no cutoff; started early, off the pair loop's path
                    double ew = 0.0, fzw = 0.0, dzw = 0.0;
                    bool near = false;
                    if (b.wall) {
                        dzw = wall_dz<false>(b, qz);
                        near = dzw * dzw < b.rc2;
                        add_zwall(b, dzw, ew, fzw);
                    }

                    // one pass: the proposed position (always) and the old one (only if it has partners)
                    unsigned hits_new, hits_old = 0;
                    if (nbm) {
                        screen_slots2<K, PZ>(b, sc, qsx, qsy, qsz, s.x[n] * b.invL, s.y[n] * b.invL, s.z[n] * b.invL,
                                         xs, ys, zs, hits_new, hits_old);
                        hits_old &= okmask;
                    } else {
                        hits_new = screen_slots<K, PZ>(b, sc, qsx, qsy, qsz, xs, ys, zs);
                    }
                    hits_new &= okmask;

                    // gas-phase fast path: nobody in range of the proposal -> all pair and site sums are exactly 0
                    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
                    unsigned in_new = 0;
                    const bool work = __any_sync(FULL, hits_new != 0) || near;
                    if (work) {
                        while (hits_new) {
                            const int k = __ffs(hits_new) - 1;
                            hits_new &= hits_new - 1;
                            const int j = particle_of(k);
                            double et, hx, hy, hz;
                            if (pair_exact(b, qx, qy, qz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz)) {
                                e += et; fx += hx; fy += hy; fz += hz;
                                in_new |= 1u << k;
                            }
                        }
                        if (near) add_sites(b, s, lane, MMpad, qx, qy, dzw, e, fx, fy, fz);
                        warp_sum4(lane, e, fx, fy, fz);
                    }
                    const double Un = 4.0 * (e + ew), Fnx = fx, Fny = fy, Fnz = fz + fzw;                 // SMC.c:319-321

                    // SMC.c:326-335: accept iff u < exp(-(Un-Um + d.(Fn+Fm)/2 + (Fn^2-Fm^2) A/(4T))/T)
                    const double Um = s.ce[n], Fmx = s.cfx[n], Fmy = s.cfy[n], Fmz = s.cfz[n];           // SMC.c:300-304, cached
                    const double f2 = fma(Fnx, Fnx, fma(Fny, Fny, Fnz * Fnz)) - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                    const double dr = fma(dX, Fnx + Fmx, fma(dY, Fny + Fmy, dZ * (Fnz + Fmz)));
                    const double xarg = -((Un - Um) + 0.5 * dr + f2 * quarterAoT) * invT;
                    const double lu = __shfl_sync(FULL, lul, t);
                    const bool acc = (lu < xarg) && (xarg > -745.1332191019411);    // exp underflows to 0 below that
                    cnt += __popc(in_new) + (lane == 0 ? nbm : 0);   // partners at the new + at the old position
                    if (acc) {
                        // partners lose the old pair terms and gain the new ones (force on j from n = -g d)
                        if (nbm) {
                            const double px = s.x[n], py = s.y[n], pz = s.z[n];
                            while (hits_old) {
                                const int k = __ffs(hits_old) - 1;
                                hits_old &= hits_old - 1;
                                const int j = particle_of(k);
                                double et, hx, hy, hz;
                                if (pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz)) {
                                    s.ce[j] -= 4.0 * et; s.cfx[j] += hx; s.cfy[j] += hy; s.cfz[j] += hz;
                                    s.nb[j] -= 1;
                                }
                            }
                        }
                        int nbn = 0;
                        if (work) {
                            unsigned hn = in_new;
                            while (hn) {
                                const int k = __ffs(hn) - 1;
                                hn &= hn - 1;
                                const int j = particle_of(k);
                                double et, hx, hy, hz;
                                pair_exact(b, qx, qy, qz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz);
                                s.ce[j] += 4.0 * et; s.cfx[j] -= hx; s.cfy[j] -= hy; s.cfz[j] -= hz;
                                s.nb[j] += 1;
                            }
                            nbn = __reduce_add_sync(FULL, __popc(in_new));
                        }
                        __syncwarp();                    // partner updates read the old position of n: order before overwriting it
                        if (lane == t) {                 // the owner: physical slot 0 is the visited slot
                            s.x[n] = qx; s.y[n] = qy; s.z[n] = qz;
                            s.ce[n] = Un; s.cfx[n] = Fnx; s.cfy[n] = Fny; s.cfz[n] = Fnz;
                            s.nb[n] = (unsigned short)nbn;
                            xs[0] = qsx; ys[0] = qsy; zs[0] = qsz;
                        }
                        E += Un - Um;                   // SMC.c:341
                        nacc++;
                    }
                    if (FED && a.accepted != nullptr && lane == 0) {
                        int nn = n - off;
                        if (nn < 0) nn += N;
                        a.accepted[sci * N + nn] = acc ? 1 : 0;
                    }
                    __syncwarp();
                }
            }
            if (seg < K) rotate();
        }
        if (a.trace_E != nullptr && lane == 0) { a.trace_E[sci] = E; a.trace_acc[sci] = nacc - nacc0; }
    }

    for (int j = lane; j < N; j += 32) {               // the shared-memory mirror holds the exact positions
        P[j] = s.x[j]; P[Npad + j] = s.y[j]; P[2 * Npad + j] = s.z[j];
    }
    if (a.cache_out) {                                  // test hook: the caches as they stand at the end
        double *co = a.cache_out + (size_t)chain * 5 * Npad;
        for (int j = lane; j < N; j += 32) {
            co[j] = s.ce[j]; co[Npad + j] = s.cfx[j]; co[2 * Npad + j] = s.cfy[j]; co[3 * Npad + j] = s.cfz[j];
            co[4 * Npad + j] = (double)s.nb[j];
        }
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
        }
    }
}


// PZ (bulk, z periodic) is a per-chain flag: one uniform branch per CTA picks the specialisation, so
// the slab-mode pair loop carries no predicated-off z-wrap instructions.
template <int K, bool FED>
__global__ void __launch_bounds__(32, (K <= 8 ? SMCB_SWEEP_MINB : 8)) k_sweep_cached(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_cached_body<K, FED, true>(d, a);
    else sweep_cached_body<K, FED, false>(d, a);
}

}  // namespace smcb
