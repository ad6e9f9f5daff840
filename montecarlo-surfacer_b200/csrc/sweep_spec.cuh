// sweep_spec.cuh — the FAST sweep kernel, segment-speculative variant (oneParticleMoves, SMC.c:278-351).
//
// Same Markov chain as k_sweep_cached (same proposal, acceptance expression, visiting order and random inputs,
// same caches, same packed-FP32 superset screen with exact FP64 terms for the hits).  k_sweep_cached resolves the
// 32 trials of a segment one after the other: screen, vote, branch, and - for every trial that finds a partner
// (~30 % of the trials of the benchmark gas) - a 560-instruction warp-wide general path for ONE pair.  That serial
// chain is what bounds it (profiles/r01: 3 warps per scheduler, 59 % issue, one trial per 1160 cycles).
//
// Here the whole segment is evaluated LANE-PARALLEL against the state at the start of the segment:
//   phase 1  lane t prepares trial t (random inputs, proposal from the cached force, flat-wall terms)    [as before]
//   phase 2  the 32 proposals are screened against all K slots in one pipelined loop (independent iterations, two
//            trials in flight).  Lane t keeps, for ITS trial: which lanes found hits (the hit masks go to a 32x32
//            table in shared memory), which of the segment's own particles are in range of its proposal (po), and
//            which of the other PROPOSALS are (pp, from a proposal-against-proposal test in the same loop, both trials
//            of an iteration in one packed pass).  Hit bits are the SIGN of r2 - rc2 moved into the mask by funnel
//            shifts (screen_slots_sgn); `in` and `pp` are collected in shift registers, one funnel shift per trial
//   phase 3  lane t evaluates the exact FP64 pair terms of its own trial's partners (0-1 of them in the gas: a
//            short divergent loop; it keeps their SUMS, up to SMCB_SPEC_MAXPARTNERS partners), adds the flat wall,
//            and decides its trial completely (SMC.c:319-335)
//   phase 4  the warp walks only the trials that need work, in visiting order: accepted ones are committed by
//            their owner lane (position, caches, the partners' caches by Newton's third law - one partner: from the
//            terms the lane holds, several: formed again, only the sums were kept); a trial whose
//            inputs were changed by an earlier accepted trial of the segment - its cached force (dirty), or the set
//            of particles in range of its proposal ((pp | po) & accepted) - or that is near the surface or has
//            more partners than that, is redone on the warp-wide general path of k_sweep_cached.  Rejected trials with
//            valid speculation cost nothing in phase 4.  What cannot change during the epochs (mine, accepted, bad,
//            lonely) is voted once per segment: an epoch is one vote and one shuffle.  A molecule with exactly one
//            partner keeps a record of who it is (p1[]), so its accepted move corrects that partner from the owner
//            lane instead of screening the old position again.
// A trial's speculation is valid exactly when nothing it read has changed since the start of the segment, so the
// results are those of the sequential sweep (tests: accept flags identical to the oracle's, positions and energies
// within 1e-12 teacher-forced; cache consistency after many sweeps).
//
// Dense states (a droplet on the wall: 30-80 partners each, every accepted move dirties a quarter of the segment)
// make the speculation worthless; a segment whose particles average more than two partners skips phases 2-3 and
// sends every trial to the general path (the behaviour of k_sweep_cached).
#pragma once

namespace smcb {

constexpr unsigned short kNoP1 = 0xffffu;
template <int K> struct HitMaskT { typedef unsigned char type; };
template <> struct HitMaskT<16> { typedef unsigned short type; };

template <int K>
struct SpecSmem {
    static constexpr int NS = 32 * K;
    static __host__ __device__ size_t bytes(int MMpad)
    {
        return ChainSmem::bytes(NS, MMpad) + (size_t)32 * 32 * sizeof(typename HitMaskT<K>::type) + 32 * sizeof(unsigned) + (size_t)NS * sizeof(unsigned short);
    }
};

#ifndef SMCB_SPEC_MINB
#define SMCB_SPEC_MINB 10
#endif
#ifndef SMCB_SPEC_MAXPARTNERS
#define SMCB_SPEC_MAXPARTNERS 3      // a trial with up to this many partners at its proposal is decided by its own lane
#endif

// screen_slots for phase 2: the hit bit of a slot is the SIGN of r2 - rc2 (one packed subtraction for two slots), shifted
// into the mask by a funnel shift - 1.5 instructions per slot instead of FSETP + predicated OR.  Slots are visited from
// the last to the first so that slot 0 ends up in bit 0.  (r2 == rc2 gives +0: no hit, like `<`.)
template <int K, bool PZ>
__device__ __forceinline__ unsigned screen_slots_sgn(const ScreenConsts &sc, float px, float py, float pz, const Slots<K> &q)
{
    const float2 ax = make_float2(px, px), ay = make_float2(py, py), az = make_float2(pz, pz);
    const float2 MG = make_float2(12582912.f, 12582912.f), rc2 = make_float2(sc.rc2s, sc.rc2s);
    unsigned hits = 0;
#pragma unroll
    for (int k = Slots<K>::KP - 1; k >= 0; k--) {
        float2 sx = sub2(ax, q.x[k]);
        sx = sub2(sx, sub2(add2(sx, MG), MG));
        float2 sy = sub2(ay, q.y[k]);
        sy = sub2(sy, sub2(add2(sy, MG), MG));
        float2 sz = sub2(az, q.z[k]);
        if (PZ) {
            const float2 t = mul2(sz, make_float2(sc.inv_zper, sc.inv_zper));
            sz = fma2(sub2(add2(t, MG), MG), make_float2(-sc.zper, -sc.zper), sz);
        }
        const float2 d = sub2(fma2(sz, sz, fma2(sy, sy, mul2(sx, sx))), rc2);
        hits = __funnelshift_l(__float_as_uint(d.y), hits, 1);
        hits = __funnelshift_l(__float_as_uint(d.x), hits, 1);
    }
    return hits;
}

template <int K, bool FED, bool PZ>
__device__ __forceinline__ void sweep_spec_body(const DevChains &d, const SweepArgs &a)
{
    typedef typename HitMaskT<K>::type HM;
    const int lane = threadIdx.x, chain = blockIdx.x;
    const int N = d.N, Npad = d.Npad;
    const int MM = d.M * d.M, MMpad = (MM + 3) & ~3;
    extern __shared__ double sm[];
    ChainSmem s;
    s.carve(sm, 32 * K, MMpad);
    unsigned *hbrow = reinterpret_cast<unsigned *>(s.site + 4 * MMpad);   // [trial]: which lanes hold hits for the trial's proposal
    HM *hm = reinterpret_cast<HM *>(hbrow + 32);               // [trial][lane]: the lane's hit mask for the trial's proposal
    // [particle]: WHO the partner is while the particle has exactly one (kNoP1: not known) - an accepted move of such a
    // particle then corrects that partner's caches from the owner lane, without screening the old position again
    unsigned short *p1 = reinterpret_cast<unsigned short *>(hm + 32 * 32);
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
        if (cntn == 1) { if (in) p1[n] = (unsigned short)(lane + 32 * (__ffs(in) - 1)); }
        else if (lane == 0) p1[n] = kNoP1;
    }
    __syncwarp();

    const double AoT = b.A / b.T;
    const double sigma = sqrt(2.0 * b.A);            // vecBoxMuller(sqrt(2.0*A), ...)  SMC.c:284
    const double quarterAoT = 0.25 * AoT, invT = 1.0 / b.T;
    double E = a.refresh_E ? Erebuilt : d.E[chain];
    double dE = 0.0;                                 // per-lane share of the running energy (owner-committed trials)
    int nacc = 0;
    unsigned cnt = 0;                                // per-lane, < 2^32 per launch
#ifdef SMCB_SPEC_STATS
    unsigned st[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};    // path statistics (profiles/spec_stats.py), warp-uniform
#define SMCB_ST(i, v) st[i] += (v)
#else
#define SMCB_ST(i, v)
#endif
    unsigned nscr = (unsigned)N;                     // N-particle screens executed (warp-uniform); the cache rebuild ran N
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};

    // physical register slot 0 always holds the slot being visited (see k_sweep_cached): `rot` = logical slot at physical 0
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
                // ================= phase 1: lane t prepares trial t (particle 32*slot+t) =================
                const int nl = 32 * slot + lane;
                const bool mine = lane >= tb && lane < te;
                __syncwarp();                           // the previous segment is done with the staging area
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, lul = 0.0;
                double p_qx = 0.0, p_qy = 0.0, p_qz = 0.0, p_ew = 0.0, p_fz = 0.0, p_dU = 0.0;
                double pe = 0.0, pgx = 0.0, pgy = 0.0, pgz = 0.0;     // pair terms of the trial's one partner at the proposal
                int pj = -1, p_np = 0, nb0 = 0;
                bool p_bad = false, p_acc = false;      // bad: near the surface or several partners -> general path
                float st_x = 0.f, st_y = 0.f, st_z = 3.0e18f;        // lanes without a trial: out of everybody's range
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
                    const double dX = fma(s.cfx[nl], AoT, g0), dY = fma(s.cfy[nl], AoT, g1), dZ = fma(s.cfz[nl], AoT, g2);   // SMC.c:307-309
                    p_qx = min_image<false>(s.x[nl] + dX, b.L, b.invL);                                    // SMC.c:311-316
                    p_qy = min_image<false>(s.y[nl] + dY, b.L, b.invL);
                    p_qz = s.z[nl] + dZ;
                    if (PZ) p_qz = min_image<false>(p_qz, b.Lz, b.invLz);
                    if (b.wall) {
                        const double dzw = wall_dz<false>(b, p_qz);
                        p_bad = dzw * dzw < b.rc2;      // surface sites in range: the general path sums them
                        add_zwall(b, dzw, p_ew, p_fz);
                    }
                    nb0 = s.nb[nl];
                    st_x = (float)(p_qx * b.invL); st_y = (float)(p_qy * b.invL); st_z = (float)(p_qz * b.invL);
                }
                s.stage[lane] = st_x; s.stage[32 + lane] = st_y; s.stage[64 + lane] = st_z;
                hbrow[lane] = 0u;
                const bool lite = __reduce_add_sync(FULL, (unsigned)nb0) > 2u * (unsigned)(te - tb);   // dense segment: no speculation
                __syncwarp();

                // per-lane interaction masks of MY trial with the other trials m of the segment:
                //   my_in  bit m: proposal m lands in range of my particle's position      (m accepted -> my caches change)
                //   my_po  bit m: particle m's position is in range of my proposal        (m accepted -> my partner set changes)
                //   my_pp  bit m: proposal m is in range of my proposal                   (m accepted -> my partner set changes)
                unsigned my_hb = 0, my_in = 0, my_po = 0, my_pp = 0;
                if (!lite) {
                    // ================= phase 2: the segment's proposals against all slots, and against each other ==========
                    // no votes in this loop: a lane records what concerns ITS particle / proposal (the relations are symmetric),
                    // and the rare hit masks go to the 32x32 table with an atomicOr on the trial's row word
                    const float MGs = 12582912.f;
                    const int t_first = tb & ~1, n_scr = (te - t_first + 1) & ~1;   // the loop below visits trials t_first .. t_first + n_scr - 1
                    nscr += (unsigned)n_scr;
                    unsigned in_sr = 0, pp_sr = 0;      // my_in / my_pp as shift registers: one funnel shift per trial, placed after the loop
                    const float2 MG2 = make_float2(MGs, MGs), rc2 = make_float2(sc.rc2s, sc.rc2s);
                    const float2 stx2 = make_float2(st_x, st_x), sty2 = make_float2(st_y, st_y), stz2 = make_float2(st_z, st_z);
                    for (int t2 = tb & ~1; t2 < te; t2 += 2) {
                        // the two proposals of this iteration, one 8-byte load per component (t2 is even)
                        const float2 ax2 = *reinterpret_cast<const float2 *>(s.stage + t2);
                        const float2 ay2 = *reinterpret_cast<const float2 *>(s.stage + 32 + t2);
                        const float2 az2 = *reinterpret_cast<const float2 *>(s.stage + 64 + t2);
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int t = t2 + h;
                            const unsigned m = screen_slots_sgn<K, PZ>(sc, h ? ax2.y : ax2.x, h ? ay2.y : ay2.x, h ? az2.y : az2.x, q) & validmask & ~((lane == t) ? 1u : 0u);
                            if (m) { hm[t * 32 + lane] = (HM)m; atomicOr(hbrow + t, 1u << lane); }
                            in_sr = __funnelshift_r(in_sr, m, 1);           // bit 0 of m enters at the top
                        }
                        // proposal against proposal, both trials of the iteration in one packed pass
                        float2 dx = sub2(ax2, stx2), dy = sub2(ay2, sty2), dz = sub2(az2, stz2);
                        dx = sub2(dx, sub2(add2(dx, MG2), MG2));
                        dy = sub2(dy, sub2(add2(dy, MG2), MG2));
                        if (PZ) dz = fma2(sub2(add2(mul2(dz, make_float2(sc.inv_zper, sc.inv_zper)), MG2), MG2), make_float2(-sc.zper, -sc.zper), dz);
                        const float2 dd = sub2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), rc2);   // sign = inside, as in screen_slots_sgn
                        pp_sr = __funnelshift_l(__float_as_uint(dd.x), pp_sr, 1);
                        pp_sr = __funnelshift_l(__float_as_uint(dd.y), pp_sr, 1);
                    }
                    if (n_scr > 0) {
                        my_in = (in_sr >> (32 - n_scr)) << t_first;                  // the first trial visited was shifted down the farthest
                        my_pp = (__brev(pp_sr) >> (32 - n_scr)) << t_first;          // ... and up the farthest here
                    }
                    my_pp &= ~(1u << lane);
                    __syncwarp();
                    my_hb = hbrow[lane];
                    // ================= phase 3: lane t evaluates and decides its own trial =================
                    if (mine && !p_bad) {
                        unsigned hb = my_hb;
                        while (hb) {
                            const int l = __ffs(hb) - 1;
                            hb &= hb - 1;
                            unsigned m = hm[lane * 32 + l];
                            my_po |= (m & 1u) << l;
                            while (m) {
                                const int k = __ffs(m) - 1;
                                m &= m - 1;
                                int sl = k + rot;
                                if (sl >= K) sl -= K;
                                const int j = l + 32 * sl;
                                double et, hx, hy, hz;
                                if (pair_exact(b, p_qx, p_qy, p_qz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz)) {
                                    pe += et; pgx += hx; pgy += hy; pgz += hz; pj = j;     // one partner: exactly its terms
                                    p_np++;
                                }
                            }
                        }
                        if (p_np > SMCB_SPEC_MAXPARTNERS) {
                            p_bad = true;
                        } else {
                            const double Um = s.ce[nl], Fmx = s.cfx[nl], Fmy = s.cfy[nl], Fmz = s.cfz[nl];   // SMC.c:300-304, cached
                            const double dX = fma(Fmx, AoT, g0), dY = fma(Fmy, AoT, g1), dZ = fma(Fmz, AoT, g2);
                            const double Un = 4.0 * (pe + p_ew), Fnz = pgz + p_fz;                            // SMC.c:319-321
                            const double f2 = fma(pgx, pgx, fma(pgy, pgy, Fnz * Fnz)) - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                            const double dr = fma(dX, pgx + Fmx, fma(dY, pgy + Fmy, dZ * (Fnz + Fmz)));
                            p_dU = Un - Um;
                            const double xarg = -(p_dU + 0.5 * dr + f2 * quarterAoT) * invT;                  // SMC.c:326-329
                            p_acc = (lul < xarg) && (xarg > -745.1332191019411);
                        }
                    }
                }

                // ================= phase 4: resolve in visiting order =================
                unsigned A = 0;                          // trials of this segment accepted so far
                SMCB_ST(0, te - tb); SMCB_ST(11, __popc(__ballot_sync(FULL, mine && nb0 > 0)));
                unsigned genmask = 0;                    // trials that went through the general path
                unsigned dirty = 0;                      // particles of the visited slot whose caches were touched by an accepted trial
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
                            s.nb[j] -= 1; p1[j] = kNoP1;          // whoever is left, it is not recorded
                            touched |= (k == 0);
                        }
                    }
                    return touched;
                };
                // A "lonely" trial found nobody at its proposal and its particle has no partners now: committing it touches
                // nobody else's caches.  Each pass of the loop is an EPOCH: f = the first pending trial that needs serial
                // work (void or possibly void speculation, or an accepted trial with partners); the accepted lonely trials
                // before f are committed together by their owner lanes, then f is handled alone.
                const bool lonely = mine && !p_bad && p_np == 0 && nb0 == 0;
                const unsigned lowmask = (1u << lane) - 1u;
                // what does not change during the epochs is voted ONCE: the loop below then needs one vote per epoch (which
                // pending trials an accepted or still open earlier trial interferes with) and one shuffle
                const unsigned mineb = __ballot_sync(FULL, mine);
                const unsigned accb = __ballot_sync(FULL, mine && p_acc);
                const unsigned badb = __ballot_sync(FULL, mine && (lite || p_bad));
                const unsigned serialb = accb & ~__ballot_sync(FULL, lonely);      // accepted, but with partners: the owner commits it alone
                int cur = tb;
                while (cur < te) {
                    const unsigned pendb = mineb & ~((1u << cur) - 1u);
                    const bool pending = (pendb >> lane) & 1u;
                    const unsigned X = my_in | my_po | my_pp;
                    // trials whose particle may still move in this segment (decision open, or accepted)
                    const unsigned move = pendb & (badb | dirty | accb);
                    if (move == 0u) {                    // the rest are rejections; void only if an ACCEPTED trial interferes
                        if (__ballot_sync(FULL, pending && (X & A) != 0u) == 0u) break;
                    }
                    const unsigned hardb = __ballot_sync(FULL, pending && (X & (A | (move & lowmask))) != 0u) | (pendb & (badb | dirty | serialb));
                    const int f = hardb ? __ffs(hardb) - 1 : 32;
                    const unsigned cm = pendb & accb & (hardb ? ((1u << f) - 1u) : 0xffffffffu);   // lonely, and nothing before them in the segment interferes
                    const bool commit = (cm >> lane) & 1u;
                    if (cm) {
                        if (commit) {
                            s.x[nl] = p_qx; s.y[nl] = p_qy; s.z[nl] = p_qz;
                            s.ce[nl] = 4.0 * p_ew; s.cfx[nl] = 0.0; s.cfy[nl] = 0.0; s.cfz[nl] = p_fz;
                            q.set(0, st_x, st_y, st_z);
                            dE += p_dU;                 // SMC.c:341, summed per lane, reduced at the end of the sweep
                        }
                        A |= cm;
                        nacc += __popc(cm);
                        SMCB_ST(1, __popc(cm));
                        __syncwarp();
                    }
                    if (f >= te) break;
                    const int t = f;
                    cur = t + 1;
                    // with the epoch's commits known: does trial t's speculation stand?
                    const bool conf_t = (((badb | dirty) >> t) & 1u) || (__shfl_sync(FULL, X, t) & A) != 0u;
                    if (!conf_t && !((accb >> t) & 1u)) { SMCB_ST(8, 1); continue; }    // a valid rejection
                    const int n = 32 * slot + t;
                    const unsigned okmask = validmask & ~((lane == t) ? 1u : 0u);
                    const int nbm = s.nb[n];
                    if (!conf_t) {
                        // ---- the speculation of trial t stands and it accepts: its owner commits it
                        SMCB_ST(2, 1); SMCB_ST(10, __shfl_sync(FULL, pj, t) >= 0 ? 1 : 0);
                        const int p1n = p1[n];
                        if (nbm == 1 && p1n != kNoP1) {  // the one old partner is on record: the owner lane corrects it
                            SMCB_ST(9, 1);
                            if (lane == t) {
                                double et, hx, hy, hz;
                                if (pair_exact(b, s.x[n], s.y[n], s.z[n], s.x[p1n], s.y[p1n], s.z[p1n], et, hx, hy, hz)) {
                                    s.ce[p1n] -= 4.0 * et; s.cfx[p1n] += hx; s.cfy[p1n] += hy; s.cfz[p1n] += hz;
                                    s.nb[p1n] -= 1; p1[p1n] = kNoP1;
                                }
                            }
                            if ((p1n >> 5) == slot) dirty |= 1u << (p1n & 31);
                        } else if (nbm) {                // the old partners forget this particle
                            SMCB_ST(9, 1);
                            __syncwarp();
                            dirty |= __ballot_sync(FULL, drop_old_partners(n, okmask));
                            __syncwarp();
                        }
                        if (lane == t) {
                            if (p_np == 1) {             // the one new partner gains the pair terms (force on j from n = -g d)
                                s.ce[pj] += 4.0 * pe; s.cfx[pj] -= pgx; s.cfy[pj] -= pgy; s.cfz[pj] -= pgz;
                                const unsigned short c = s.nb[pj] + 1;
                                s.nb[pj] = c; p1[pj] = (c == 1) ? (unsigned short)n : kNoP1;
                            } else if (p_np > 1) {       // several: the lane holds their SUMS; each one's terms are formed again
                                unsigned hb = my_hb;
                                while (hb) {
                                    const int l = __ffs(hb) - 1;
                                    hb &= hb - 1;
                                    unsigned m = hm[lane * 32 + l];
                                    while (m) {
                                        const int k = __ffs(m) - 1;
                                        m &= m - 1;
                                        int sl = k + rot;
                                        if (sl >= K) sl -= K;
                                        const int j = l + 32 * sl;
                                        double et, hx, hy, hz;
                                        if (pair_exact(b, p_qx, p_qy, p_qz, s.x[j], s.y[j], s.z[j], et, hx, hy, hz)) {
                                            s.ce[j] += 4.0 * et; s.cfx[j] -= hx; s.cfy[j] -= hy; s.cfz[j] -= hz;
                                            const unsigned short c = s.nb[j] + 1;
                                            s.nb[j] = c; p1[j] = (c == 1) ? (unsigned short)n : kNoP1;
                                        }
                                    }
                                }
                            }
                            s.x[n] = p_qx; s.y[n] = p_qy; s.z[n] = p_qz;
                            s.ce[n] = 4.0 * (pe + p_ew); s.cfx[n] = pgx; s.cfy[n] = pgy; s.cfz[n] = pgz + p_fz;
                            s.nb[n] = (unsigned short)p_np; p1[n] = (p_np == 1) ? (unsigned short)pj : kNoP1;
                            q.set(0, st_x, st_y, st_z);
                            dE += p_dU;                 // SMC.c:341, summed per lane, reduced at the end of the sweep
                        }
                        // partners that belong to the visited slot have had their caches touched: their speculation is void.
                        // One partner: exactly that one; several: every slot-0 molecule the screen found near the proposal
                        // (a superset - a void speculation only costs the general path)
                        const int pjt = __shfl_sync(FULL, pj, t), npt = __shfl_sync(FULL, p_np, t);
                        const unsigned pot = __shfl_sync(FULL, my_po, t);
                        if (npt == 1) { if ((pjt >> 5) == slot) dirty |= 1u << (pjt & 31); }
                        else if (npt > 1) dirty |= pot;
                        A |= 1u << t;
                        nacc++;
                        __syncwarp();
                        continue;
                    }
                    // ---- general path: trial t is evaluated by the whole warp against the CURRENT state
                    genmask |= 1u << t;
#ifdef SMCB_SPEC_STATS
                    {
                        const bool bad_t = __shfl_sync(FULL, (int)p_bad, t) != 0, multi_t = __shfl_sync(FULL, p_np, t) > 1;
                        st[3]++;
                        if (lite) st[4]++;
                        else if (bad_t && multi_t) st[5]++;
                        else if (bad_t) st[6]++;
                        else if ((dirty >> t) & 1u) st[7]++;
                    }
#endif
                    const bool reuse = !((dirty >> t) & 1u);        // its cached force is untouched: the proposal of phase 1 stands
                    __syncwarp();
                    const double dX = fma(s.cfx[n], AoT, __shfl_sync(FULL, g0, t));
                    const double dY = fma(s.cfy[n], AoT, __shfl_sync(FULL, g1, t));
                    const double dZ = fma(s.cfz[n], AoT, __shfl_sync(FULL, g2, t));
                    double qx, qy, qz, ew = 0.0, fzw = 0.0, dzw = 0.0;
                    float qsx, qsy, qsz;
                    bool near = false;
                    if (reuse) {
                        qx = __shfl_sync(FULL, p_qx, t); qy = __shfl_sync(FULL, p_qy, t); qz = __shfl_sync(FULL, p_qz, t);
                        ew = __shfl_sync(FULL, p_ew, t); fzw = __shfl_sync(FULL, p_fz, t);
                        qsx = s.stage[t]; qsy = s.stage[32 + t]; qsz = s.stage[64 + t];
                        if (b.wall) { dzw = wall_dz<false>(b, qz); near = dzw * dzw < b.rc2; }
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
                    }
                    unsigned hits_new = screen_slots<K, PZ>(sc, qsx, qsy, qsz, q) & okmask;
                    nscr++;

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
                        // fetched with independent shuffles instead of the butterfly; lane order is fixed either way.
                        unsigned holders = __ballot_sync(FULL, in_new != 0 || (near && lane < MM));
                        if (__popc(holders) <= 2) {
                            double te_ = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;
                            while (holders) {
                                const int src = __ffs(holders) - 1;
                                holders &= holders - 1;
                                te_ += __shfl_sync(FULL, e, src); tx += __shfl_sync(FULL, fx, src);
                                ty += __shfl_sync(FULL, fy, src); tz += __shfl_sync(FULL, fz, src);
                            }
                            e = te_; fx = tx; fy = ty; fz = tz;
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
                        int nbn = 0, j1 = kNoP1;
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
                                const unsigned short c = s.nb[j] + 1;
                                s.nb[j] = c; p1[j] = (c == 1) ? (unsigned short)n : kNoP1;
                                touched |= (k == 0);
                                j1 = j;
                            }
                            nbn = __reduce_add_sync(FULL, __popc(in_new));
                            if (nbn == 1) j1 = __shfl_sync(FULL, j1, __ffs(__ballot_sync(FULL, in_new != 0)) - 1);   // the one new partner, for the record
                        }
                        dirty |= __ballot_sync(FULL, touched);
                        __syncwarp();                    // partner updates read the old position of n: order before overwriting it
                        if (lane == t) {                 // the owner: physical slot 0 is the visited slot
                            s.x[n] = qx; s.y[n] = qy; s.z[n] = qz;
                            s.ce[n] = Un; s.cfx[n] = Fnx; s.cfy[n] = Fny; s.cfz[n] = Fnz;
                            s.nb[n] = (unsigned short)nbn; p1[n] = (nbn == 1) ? (unsigned short)j1 : kNoP1;
                            q.set(0, qsx, qsy, qsz);
                        }
                        E += Un - Um;                   // SMC.c:341
                        nacc++;
                        A |= 1u << t;
                        if (!lite && pending && lane > t) {   // pending proposals against the position n actually moved to
                            float dx = st_x - qsx, dy = st_y - qsy, dz = st_z - qsz;
                            const float MGs = 12582912.f;
                            dx -= (dx + MGs) - MGs;
                            dy -= (dy + MGs) - MGs;
                            if (PZ) dz = fmaf((dz * sc.inv_zper + MGs) - MGs, -sc.zper, dz);
                            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            my_pp = (r2 < sc.rc2s) ? (my_pp | (1u << t)) : (my_pp & ~(1u << t));
                        }
                    }
                    __syncwarp();
                }
                // in-cutoff pair statistics of the trials that never reached the general path (the reference evaluates them)
                if (mine && !((genmask >> lane) & 1u)) cnt += (unsigned)(p_np + nb0);
                if (FED && a.accepted != nullptr && mine) {
                    int nn = nl - off;
                    if (nn < 0) nn += N;
                    a.accepted[sci * N + nn] = (unsigned char)((A >> lane) & 1u);
                }
            }
            if (seg < K) rotate();
        }
        E += warp_sum(dE);
        dE = 0.0;
        if (a.trace_E != nullptr && lane == 0) { a.trace_E[sci] = E; a.trace_acc[sci] = nacc - nacc0; }
    }

    __syncwarp();
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
#ifdef SMCB_SPEC_STATS
            for (int i = 0; i < 12; i++) atomicAdd(d.pair_counts + 3 + i, (unsigned long long)st[i]);
#endif
        }
    }
}

#ifdef SMCB_SPEC_MAXREG
#define SMCB_SPEC_BOUNDS __maxnreg__(SMCB_SPEC_MAXREG)
#else
#define SMCB_SPEC_BOUNDS __launch_bounds__(32, (K <= 8 ? SMCB_SPEC_MINB : 8))
#endif
template <int K, bool FED>
__global__ void SMCB_SPEC_BOUNDS k_sweep_spec(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_spec_body<K, FED, true>(d, a);
    else sweep_spec_body<K, FED, false>(d, a);
}

}  // namespace smcb
