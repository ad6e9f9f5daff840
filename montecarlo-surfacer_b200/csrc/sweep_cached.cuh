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
    float *stage;             // [3][32]: speculative proposals of the slot being visited, in box units (screen precision)
    // NS = 32*K is a compile-time constant of the kernel instance, so every array but `site` sits at a
    // constant offset from one base register
    __device__ __forceinline__ void carve(double *base, int NS, int MMpad)
    {
        x = base; y = x + NS; z = y + NS;
        ce = z + NS; cfx = ce + NS; cfy = cfx + NS; cfz = cfy + NS;
        stage = reinterpret_cast<float *>(cfz + NS);
        nb = reinterpret_cast<unsigned short *>(stage + 3 * 32);
        site = reinterpret_cast<double *>(nb + NS);       // NS is a multiple of 32: 8-byte aligned
    }
    static __host__ __device__ size_t bytes(int NS, int MMpad)
    {
        return (size_t)(7 * NS + 4 * MMpad) * sizeof(double) + (size_t)(3 * 32) * sizeof(float) + (size_t)NS * sizeof(unsigned short);
    }
};

// ---- phase 1: the screen, in packed single precision ---------------------------------------------
// The screen only has to find a SUPERSET of the partners inside the cutoff: every hit is re-tested and
// evaluated in double precision from the exact positions (pair_exact), and a pair outside the cutoff
// contributes exactly 0, so the results do not depend on the screen's precision.  On B200 the FP64 pipe
// issues one warp instruction per 2 cycles and is the scarce resource (profiles/r01: screen_thr.txt):
// an 8-slot FP64 screen costs ~224 SM-sub-partition cycles, the packed FP32 one ~130.  Positions are kept
// in box units (x/L) as float2 pairs of slots; a minimum image is s - rint(s) with the 1.5*2^23 trick, and
// add/sub/mul/fma.f32x2 (FADD2/FMUL2/FFMA2, sm_100+) work on two slots per instruction.
struct ScreenConsts {
    float rc2s;               // (rc/L + margin)^2: cutoff in box units, inflated by the FP32 error bound below
    float zper, inv_zper;     // Lz/L and L/Lz (bulk mode only)
    float interior;           // a point with |x|/L below this needs no minimum image along x (same for y), see below
};

// FP32 error bound of a box-unit separation: operands are rounded to float, differences and the wrap add one more
// rounding each.  Four times that bound is added to the cutoff RADIUS, so a pair inside the true cutoff can never be
// screened out.  The operands' magnitude comes from the chain's EXTENT - the largest |x|/L, |y|/L and |z|/L the engine
// found when the positions were uploaded (k_chain_extent): callers may hand over configurations that are not wrapped
// into the primary cell (the reference tolerates that too, it only wraps the molecule it moves, SMC.c:315-316).
// Moves wrap x,y into [-L/2, L/2] and the walls bound z, so the extent at upload bounds every later position.
__device__ __forceinline__ ScreenConsts make_screen(const Box &b, const float *extent = nullptr)
{
    const double eps = 1.1920929e-7;                       // 2^-23
    const double axy = fmax(0.5, extent ? (double)extent[0] : 0.5);
    const double zmax = fmax(b.Lz * b.invL + 1.0, extent ? (double)extent[1] : 0.0);
    const double delta = eps * (4.0 * axy + 2.0 * zmax);
    const double rcs = sqrt(b.rc2) * b.invL + 4.0 * delta;
    ScreenConsts sc;
    sc.rc2s = (float)(rcs * rcs * (1.0 + 1e-6));
    sc.zper = (float)(b.Lz * b.invL); sc.inv_zper = (float)(b.L * b.invLz);
    // A point farther than the (inflated) cutoff from both periodic faces along an axis: every molecule of the primary
    // cell (|x|/L <= 1/2) whose minimum-image separation along that axis is inside the cutoff has it as its PLAIN
    // difference (|plain| > 1/2 means the image is at least 1 - (|point| + 1/2) >= cutoff away), and a plain difference
    // is never shorter than the minimum image.  The screen may then skip the wrap of that axis for the whole pass.
    // Only when the chain was uploaded inside the primary cell (extent 1/2); otherwise no point qualifies.
    sc.interior = (axy <= 0.5) ? (float)(0.5 - rcs * (1.0 + 1e-6) - 4.0 * eps) : -1.f;
    return sc;
}

__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

// the lane's K particles (slots), box units, packed two slots per float2; odd K pads with a far-away slot
template <int K>
struct Slots {
    static constexpr int KP = (K + 1) / 2;
    float2 x[KP], y[KP], z[KP];
    __device__ __forceinline__ void set(int k, float X, float Y, float Z)       // k is a compile-time constant at every call site
    {
        if (k & 1) { x[k >> 1].y = X; y[k >> 1].y = Y; z[k >> 1].y = Z; }
        else       { x[k >> 1].x = X; y[k >> 1].x = Y; z[k >> 1].x = Z; }
    }
    __device__ __forceinline__ void get(int k, float &X, float &Y, float &Z) const
    {
        if (k & 1) { X = x[k >> 1].y; Y = y[k >> 1].y; Z = z[k >> 1].y; }
        else       { X = x[k >> 1].x; Y = y[k >> 1].x; Z = z[k >> 1].x; }
    }
};

// which of the lane's K slots are within the (inflated) cutoff of the point (px,py,pz), box units.
// WX / WY = false: the caller knows the point is interior along that axis (ScreenConsts::interior), no wrap needed.
template <int K, bool PZ, bool WX = true, bool WY = true>
__device__ __forceinline__ unsigned screen_slots(const ScreenConsts &sc, float px, float py, float pz, const Slots<K> &q)
{
    const float2 ax = make_float2(px, px), ay = make_float2(py, py), az = make_float2(pz, pz);
    const float2 MG = make_float2(12582912.f, 12582912.f);           // 1.5 * 2^23: (s + MG) - MG = rint(s)
    unsigned hits = 0;
#pragma unroll
    for (int k = 0; k < Slots<K>::KP; k++) {
        float2 sx = sub2(ax, q.x[k]);
        if (WX) sx = sub2(sx, sub2(add2(sx, MG), MG));
        float2 sy = sub2(ay, q.y[k]);
        if (WY) sy = sub2(sy, sub2(add2(sy, MG), MG));
        float2 sz = sub2(az, q.z[k]);
        if (PZ) {
            const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
            sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
        }
        const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
        if (r2.x < sc.rc2s) hits |= 1u << (2 * k);
        if (r2.y < sc.rc2s) hits |= 2u << (2 * k);
    }
    return hits;
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

// pair_exact without the early exit: the 12-6 terms are formed unconditionally and zeroed by selects when the pair is
// outside the true cutoff (adding 0.0 changes no sum), so two of them can be in flight at once - the hit loops below
// of allparticle_fast.cuh take two partners per iteration; their dependent FP64 chains (~250 cycles each) bound the
// condensed phase.  (In the sweep kernel, where a lane usually holds ONE partner, the second evaluation only costs: tried, slower.)
__device__ __forceinline__ bool pair_terms_nb(const Box &b, double px, double py, double pz, double jx, double jy, double jz,
                                              double &e, double &gx, double &gy, double &gz)
{
    double dx, dy, dz;
    const double r2 = pair_sep<false>(b, px, py, pz, jx, jy, jz, dx, dy, dz);
    const bool in = r2 < b.rc2;
    const double i2 = fast_rcp(in ? r2 : 1.0);
    const double i6 = i2 * i2 * i2;
    const double et = fma(i6, i6, -i6);
    const double g = i2 * i6 * fma(48.0, i6, -24.0);
    e = in ? et : 0.0;
    gx = in ? g * dx : 0.0; gy = in ? g * dy : 0.0; gz = in ? g * dz : 0.0;
    return in;
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
                                           unsigned okmask, double px, double py, double pz, const Slots<K> &q,
                                           double &U, double &Fx, double &Fy, double &Fz, unsigned &in, double *Upair = nullptr)
{
    unsigned hits = screen_slots<K, PZ>(sc, (float)(px * b.invL), (float)(py * b.invL), (float)(pz * b.invL), q) & okmask;
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    in = add_hits(b, s, lane, hits, px, py, pz, e, fx, fy, fz);
    if (Upair) *Upair = 4.0 * warp_sum(e);          // energySingle alone: the chain energy counts every pair once (SMC.c:626-646)
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
    s.carve(sm, 32 * K, MMpad);
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const double *W = d.W + (size_t)cp.wall * 2 * MM;
    double *P = d.pos + (size_t)chain * 3 * Npad;

    Slots<K> q;
    if (K & 1) q.set(K, 0.f, 0.f, 3.0e18f);              // pad slot of an odd K: never within the cutoff
    unsigned validmask = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        const bool in = j < N;
        const double X = in ? P[j] : 0.0, Y = in ? P[Npad + j] : 0.0, Z = in ? P[2 * Npad + j] : 0.0;
        if (j < Npad) { s.x[j] = X; s.y[j] = Y; s.z[j] = Z; }
        q.set(k, (float)(X * b.invL), (float)(Y * b.invL), (float)(Z * b.invL));
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
    const ScreenConsts sc = make_screen(b, d.extent ? d.extent + 2 * chain : nullptr);

    // ---- rebuild the caches from the positions ------------------------------------
    double Erebuilt = 0.0;
    for (int n = 0; n < N; n++) {
        const unsigned okmask = validmask & ~(((n & 31) == lane) ? (1u << (n >> 5)) : 0u);
        double U, Fx, Fy, Fz, Up;
        unsigned in;
        eval_point<K, PZ>(b, sc, s, lane, MMpad, okmask, s.x[n], s.y[n], s.z[n], q, U, Fx, Fy, Fz, in, &Up);
        Erebuilt += U - 0.5 * Up;                    // energy(R) + wallsEnergy(R), SMC.c:48
        const int cntn = __reduce_add_sync(FULL, __popc(in));
        if (lane == 0) { s.ce[n] = U; s.cfx[n] = Fx; s.cfy[n] = Fy; s.cfz[n] = Fz; s.nb[n] = (unsigned short)cntn; }
    }
    __syncwarp();

    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    const double quarterAoT = 0.25 * AoT, invT = 1.0 / b.T;
    double E = a.refresh_E ? Erebuilt : d.E[chain];
    double dE = 0.0;                                 // per-lane share of the running energy (speculative path)
    int nacc = 0;
    unsigned cnt = 0;                                // per-lane, < 2^32 per launch
    unsigned nscr = (unsigned)N;                     // N-particle screens executed (warp-uniform); the cache rebuild ran N
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};

    // Physical register slot 0 always holds the slot that is being visited: the visiting
    // order n = (nn+offset)%N walks the slots cyclically, so the register arrays are rotated
    // by one at every slot boundary (48 moves per 32 trials) instead of being updated through
    // a run-time index at every accepted trial.  `rot` = logical slot held at physical 0.
    int rot = 0;
    auto rotate = [&]() {
        if (K > 1) {
            float tx, ty, tz, ux, uy, uz;
            q.get(0, tx, ty, tz);
#pragma unroll
            for (int k = 0; k + 1 < K; k++) { q.get(k + 1, ux, uy, uz); q.set(k, ux, uy, uz); }
            q.set(K - 1, tx, ty, tz);
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
                // Each lane prepares its own particle of this slot (lane <-> particle 32*slot+lane): the random
                // inputs, and SPECULATIVELY the whole trial under the assumption that nobody is in range of the
                // proposal - proposal from the cached force, flat-wall terms there, acceptance.  Then the trials
                // are resolved in visiting order, each on the cheapest path that is still exact:
                //   fast    the screen finds nobody near the proposal: the speculated decision stands; if the move
                //           is accepted and the OLD position had partners, they lose the pair terms (medium);
                //   general partners at the proposal, or the proposal is within the cutoff of the surface, or an
                //           earlier accepted trial of this segment touched this particle's caches (dirty):
                //           energy/force at the proposal are summed over the warp; the speculated proposal and
                //           flat-wall terms are reused unless the speculation is void (near / dirty).
                const int nl = 32 * slot + lane;
                const bool mine = lane >= tb && lane < te;
                __syncwarp();                           // the previous segment is done with the staging area
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, lul = 0.0;
                double p_qx = 0.0, p_qy = 0.0, p_qz = 0.0, p_ew = 0.0, p_fz = 0.0, p_dU = 0.0;
                bool p_near = false, p_acc = false;
                if (mine) {
                    double ul;
                    if (FED) {
                        const double *dsp = a.displ + sci * 3 * N;
                        g0 = dsp[3 * nl]; g1 = dsp[3 * nl + 1]; g2 = dsp[3 * nl + 2];
                        int nn = nl - off;              // trial index in visiting order (u is per trial, SMC.c:335)
                        if (nn < 0) nn += N;
                        ul = a.u[sci * N + nn];
                    } else {
                        rng_particle_gauss_f32(id, step, (uint32_t)nl, g0, g1, g2);
                        g0 *= sigma; g1 *= sigma; g2 *= sigma;
                        ul = rng_particle_uniform(id, step, (uint32_t)nl);
                    }
                    lul = log(ul);                      // u < exp(x)  <=>  log(u) < x, evaluated lane-parallel
                    const double Fmx = s.cfx[nl], Fmy = s.cfy[nl], Fmz = s.cfz[nl], Um = s.ce[nl];
                    const double dX = fma(Fmx, AoT, g0), dY = fma(Fmy, AoT, g1), dZ = fma(Fmz, AoT, g2);   // SMC.c:307-309
                    p_qx = min_image<false>(s.x[nl] + dX, b.L, b.invL);                                    // SMC.c:311-316
                    p_qy = min_image<false>(s.y[nl] + dY, b.L, b.invL);
                    p_qz = s.z[nl] + dZ;
                    if (PZ) p_qz = min_image<false>(p_qz, b.Lz, b.invLz);
                    if (b.wall) {
                        const double dzw = wall_dz<false>(b, p_qz);
                        p_near = dzw * dzw < b.rc2;
                        add_zwall(b, dzw, p_ew, p_fz);
                    }
                    const double Un0 = 4.0 * p_ew;                                                        // SMC.c:319, no partners
                    const double f2 = p_fz * p_fz - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                    const double dr = fma(dX, Fmx, fma(dY, Fmy, dZ * (p_fz + Fmz)));
                    p_dU = Un0 - Um;
                    const double xarg = -(p_dU + 0.5 * dr + f2 * quarterAoT) * invT;                      // SMC.c:326-329
                    p_acc = (lul < xarg) && (xarg > -745.1332191019411);
                    s.stage[lane] = (float)(p_qx * b.invL); s.stage[32 + lane] = (float)(p_qy * b.invL); s.stage[64 + lane] = (float)(p_qz * b.invL);
                }
                __syncwarp();
                const unsigned near0 = __ballot_sync(FULL, p_near);
                const unsigned acc0 = __ballot_sync(FULL, p_acc);
                unsigned dirty = 0;
                // the partners of particle m's CURRENT (old) position lose their pair terms with m (force on j from m = -g d);
                // returns whether a particle of the slot being visited (physical slot 0) was touched
                auto drop_old_partners = [&](int m, unsigned okm) -> bool {
                    const double px = s.x[m], py = s.y[m], pz = s.z[m];
                    unsigned ho = screen_slots<K, PZ>(sc, (float)(px * b.invL), (float)(py * b.invL), (float)(pz * b.invL), q) & okm;
                    nscr++;
                    bool touched = false;
                    while (ho) {
                        const int k = __ffs(ho) - 1;
                        ho &= ho - 1;
                        const int j = particle_of(k);
                        double et, hx, hy, hz;
                        if (pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz)) {
                            s.ce[j] -= 4.0 * et; s.cfx[j] += hx; s.cfy[j] += hy; s.cfz[j] += hz;
                            s.nb[j] -= 1;
                            touched |= (k == 0);
                        }
                    }
                    return touched;
                };
                for (int t = tb; t < te; t++) {
                    const int n = 32 * slot + t;
                    const unsigned okmask = validmask & ~((lane == t) ? 1u : 0u);
                    const bool spec = !(((near0 | dirty) >> t) & 1u);       // lane t's speculation stands
                    const int nbm = s.nb[n];
                    unsigned hits_new = 0;
                    float qsx = 0.f, qsy = 0.f, qsz = 0.f;
                    if (spec) {
                        // ---- one screen of the lane's K slots against the staged proposal
                        qsx = s.stage[t]; qsy = s.stage[32 + t]; qsz = s.stage[64 + t];
                        hits_new = screen_slots<K, PZ>(sc, qsx, qsy, qsz, q) & okmask;
                        nscr++;
                        if (!__any_sync(FULL, hits_new != 0)) {
                            const bool facc = (acc0 >> t) & 1u;
                            if (lane == 0) cnt += nbm;   // in-cutoff pairs of the old position (the reference evaluates them)
                            if (facc) {
                                if (nbm) {               // medium path: the old partners forget this particle
                                    __syncwarp();
                                    dirty |= __ballot_sync(FULL, drop_old_partners(n, okmask));
                                    __syncwarp();
                                }
                                if (lane == t) {         // the owner: its registers hold the proposal and its energy/force
                                    s.x[n] = p_qx; s.y[n] = p_qy; s.z[n] = p_qz;
                                    s.ce[n] = 4.0 * p_ew; s.cfx[n] = 0.0; s.cfy[n] = 0.0; s.cfz[n] = p_fz;
                                    s.nb[n] = 0;
                                    q.set(0, qsx, qsy, qsz);
                                    dE += p_dU;         // SMC.c:341, summed per lane, reduced at the end of the sweep
                                }
                                nacc++;
                            }
                            if (FED && a.accepted != nullptr && lane == 0) {
                                int nn = n - off;
                                if (nn < 0) nn += N;
                                a.accepted[sci * N + nn] = facc ? 1 : 0;
                            }
                            continue;
                        }
                    }
                    // ---- general path
                    __syncwarp();
                    // proposal from the cached force of particle n (SMC.c:303-316)
                    const double dX = fma(s.cfx[n], AoT, __shfl_sync(FULL, g0, t));
                    const double dY = fma(s.cfy[n], AoT, __shfl_sync(FULL, g1, t));
                    const double dZ = fma(s.cfz[n], AoT, __shfl_sync(FULL, g2, t));
                    double qx, qy, qz, ew = 0.0, fzw = 0.0, dzw = 0.0;
                    bool near = false;
                    if (spec) {                          // as lane t speculated them: same values, no recomputation
                        qx = __shfl_sync(FULL, p_qx, t); qy = __shfl_sync(FULL, p_qy, t); qz = __shfl_sync(FULL, p_qz, t);
                        ew = __shfl_sync(FULL, p_ew, t); fzw = __shfl_sync(FULL, p_fz, t);
                    } else {
                        qx = min_image<false>(s.x[n] + dX, b.L, b.invL);
                        qy = min_image<false>(s.y[n] + dY, b.L, b.invL);
                        qz = s.z[n] + dZ;
                        if (PZ) qz = min_image<false>(qz, b.Lz, b.invLz);
                        qsx = (float)(qx * b.invL); qsy = (float)(qy * b.invL); qsz = (float)(qz * b.invL);
                        if (b.wall) {                    // flat wall at the proposal: uniform, no cutoff
                            dzw = wall_dz<false>(b, qz);
                            near = dzw * dzw < b.rc2;
                            add_zwall(b, dzw, ew, fzw);
                        }
                        hits_new = screen_slots<K, PZ>(sc, qsx, qsy, qsz, q) & okmask;
                        nscr++;
                    }

                    // gas-phase fast path: nobody in range of the proposal -> all pair and site sums are exactly 0
                    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
                    double le = 0.0, lx = 0.0, ly = 0.0, lz = 0.0;
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
                        // a lane with exactly one partner (the usual case) holds that pair's terms in e, fx, fy, fz:
                        // keep them for the partner's cache update instead of evaluating the pair again
                        le = e; lx = fx; ly = fy; lz = fz;
                        if (near) add_sites(b, s, lane, MMpad, qx, qy, dzw, e, fx, fy, fz);
                        // Sum over the warp.  Usually one or two lanes hold a partner: their four partial sums are
                        // fetched with independent shuffles (one shuffle latency) instead of the five dependent
                        // levels of the butterfly; lane order is fixed, so the result is deterministic either way.
                        unsigned holders = __ballot_sync(FULL, in_new != 0 || (near && lane < MM));
                        if (__popc(holders) <= 2) {
                            double te = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;
                            while (holders) {
                                const int src = __ffs(holders) - 1;
                                holders &= holders - 1;
                                te += __shfl_sync(FULL, e, src); tx += __shfl_sync(FULL, fx, src);
                                ty += __shfl_sync(FULL, fy, src); tz += __shfl_sync(FULL, fz, src);
                            }
                            e = te; fx = tx; fy = ty; fz = tz;
                        } else {
                            warp_sum4(lane, e, fx, fy, fz);
                        }
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
                        bool touched = false;           // physical slot 0 = the slot being visited: its speculation is void
                        if (nbm) touched = drop_old_partners(n, okmask);
                        int nbn = 0;
                        if (work) {
                            unsigned hn = in_new;
                            const bool single = __popc(in_new) == 1 && !(near && lane < MM);   // le.. are that one pair's terms
                            while (hn) {
                                const int k = __ffs(hn) - 1;
                                hn &= hn - 1;
                                const int j = particle_of(k);
                                double et = le, hx = lx, hy = ly, hz = lz;
                                if (!single) pair_exact(b, qx, qy, qz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz);
                                s.ce[j] += 4.0 * et; s.cfx[j] -= hx; s.cfy[j] -= hy; s.cfz[j] -= hz;
                                s.nb[j] += 1;
                                touched |= (k == 0);
                            }
                            nbn = __reduce_add_sync(FULL, __popc(in_new));
                        }
                        dirty |= __ballot_sync(FULL, touched);
                        __syncwarp();                    // partner updates read the old position of n: order before overwriting it
                        if (lane == t) {                 // the owner: physical slot 0 is the visited slot
                            s.x[n] = qx; s.y[n] = qy; s.z[n] = qz;
                            s.ce[n] = Un; s.cfx[n] = Fnx; s.cfy[n] = Fny; s.cfz[n] = Fnz;
                            s.nb[n] = (unsigned short)nbn;
                            q.set(0, qsx, qsy, qsz);
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
        E += warp_sum(dE);
        dE = 0.0;
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
            atomicAdd(d.pair_counts + 2, (unsigned long long)nscr * (unsigned long long)(N - 1));
        }
    }
}


// PZ (bulk, z periodic) is a per-chain flag: one uniform branch per CTA picks the specialisation, so
// the slab-mode pair loop carries no predicated-off z-wrap instructions.
#ifdef SMCB_SWEEP_MAXREG
#define SMCB_SWEEP_BOUNDS __maxnreg__(SMCB_SWEEP_MAXREG)
#else
#define SMCB_SWEEP_BOUNDS __launch_bounds__(32, (K <= 8 ? SMCB_SWEEP_MINB : 8))
#endif
template <int K, bool FED>
__global__ void SMCB_SWEEP_BOUNDS k_sweep_cached(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_cached_body<K, FED, true>(d, a);
    else sweep_cached_body<K, FED, false>(d, a);
}

}  // namespace smcb
