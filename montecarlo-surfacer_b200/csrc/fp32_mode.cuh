// fp32_mode.cuh — the optional single-precision mode (SMCB_FP32) of the static evaluation and of the all-particle step.
//
// north_star: "energies and forces ... within 1e-12 relative error in fp64 (1e-5 relative in the optional fp32 mode)".
// Everything between the load of a configuration and the final block reduction is FP32 arithmetic - no F2F in the
// pair loop (round 1 measured mixed FP64-separation / FP32-terms arithmetic 15-20 % SLOWER than all-FP64 on B200).
// What keeps 1e-5 reachable with coordinates up to |z| = Lz/2 = 120 (float spacing there: 8e-6, already 3e-6 of a
// separation at the cutoff) is the REPRESENTATION, not the arithmetic: a coordinate is carried as an unevaluated pair
// of floats hi + lo (hi = fl(x), lo = fl(x - hi); 48 bits), and a separation is (hi_i - hi_j) + (lo_i - lo_j) - the
// first difference is exact for molecules a few sigma apart, so separations are good to 1e-7 of THEMSELVES wherever
// the molecules sit in the box.  The 12-6 terms then cost ~6 roundings: 5e-7.  The box lengths are pairs too, so the
// minimum image does not inherit the rounding of L.
//
// Same physics as the FP64 kernels (SMC.c:557-895 restated: pair energy / force / virial with minimum image in x,y,
// flat wall + sites, wallsPressure as the reference writes it and as it meant it), same Philox stream as the FAST
// all-particle kernel, same acceptance expression (SMC.c:326-329 summed over the molecules).  The Metropolis sums are
// accumulated per thread in FP32 and reduced over the block in FP64.  Limits that are the format's, not the code's:
// a molecule beyond a wall gets the reference's 1e-4 clamp, whose r^-12 overflows a float (inf: the move is rejected,
// the energy of such a START configuration is not representable); a chain's total energy carries ~1e-6 of sum |e_ij|.
// Not the headline path: bench.py reports it on its own line with "dtype": "f32".
#pragma once

namespace smcb {

struct FF { float hi, lo; };                                    // value = hi + lo, |lo| <= ulp(hi)/2

__device__ __forceinline__ FF ff_split(double x)
{
    FF r;
    r.hi = (float)x;
    r.lo = (float)(x - (double)r.hi);
    return r;
}

// hi + lo + d, renormalised (Knuth two-sum on the leading parts)
__device__ __forceinline__ FF ff_add(FF a, float d)
{
    const float s = a.hi + d;
    const float bb = s - a.hi;
    const float err = (a.hi - (s - bb)) + (d - bb);
    const float lo = a.lo + err;
    FF r;
    r.hi = s + lo;
    r.lo = lo - (r.hi - s);
    return r;
}

struct BoxF {
    FF L, Lz;
    float invL, invLz, rc2, a0, b0;
    bool wall;
    int M;
};

__device__ __forceinline__ BoxF make_boxf(const Box &b)
{
    BoxF f;
    f.L = ff_split(b.L); f.Lz = ff_split(b.Lz);
    f.invL = (float)b.invL; f.invLz = (float)b.invLz;
    f.rc2 = (float)b.rc2; f.a0 = (float)b.a0; f.b0 = (float)b.b0;
    f.wall = b.wall; f.M = b.M;
    return f;
}

// d - P*rint(d/P) with P = hi + lo: k is a small integer, k*P.hi and k*P.lo are single roundings folded by FMAs
__device__ __forceinline__ float wrapf(float d, FF P, float invP)
{
    const float k = rintf(d * invP);
    return fmaf(-k, P.lo, fmaf(-k, P.hi, d));
}

// separation a - b of two hi+lo coordinates
__device__ __forceinline__ float sepf(float ah, float al, float bh, float bl) { return (ah - bh) + (al - bl); }

// minimum-image separation a - b - k P.  ah - bh is exact for neighbours on the same side of the box, but a pair that is
// close ACROSS the periodic boundary has |ah - bh| ~ L: that difference rounds at ulp(L) (4e-6 for L = 33) before the
// wrap brings it down to ~1 - the rounding error (recovered exactly, two-sum) is added back in that rare case.
__device__ __forceinline__ float sepwrapf(float ah, float al, float bh, float bl, FF P, float invP)
{
    const float s = ah - bh;
    const float k = rintf(s * invP);
    float d = fmaf(-k, P.hi, s) + fmaf(-k, P.lo, al - bl);
    if (k != 0.f) {
        const float bb = s - ah;
        d += (ah - (s - bb)) - (bh + bb);
    }
    return d;
}

__device__ __forceinline__ float rcpf(float x) { return __frcp_rn(x); }      // correctly rounded 1/x (MUFU.RCP + fix-up)

// shared-memory image of one configuration: x,y,z as hi/lo pairs
struct ConfF {
    float *xh, *xl, *yh, *yl, *zh, *zl;
    __device__ __forceinline__ void carve(float *base, int Npad)
    {
        xh = base; xl = xh + Npad; yh = xl + Npad; yl = yh + Npad; zh = yl + Npad; zl = zh + Npad;
    }
};

// molecule at (x,y,z) against every other molecule of the staged configuration: LJ energy (already *4), force, virial
// sum and in-cutoff count, all FP32
template <bool PZ>
__device__ __forceinline__ void pairs_f32(const BoxF &b, const ConfF &c, int N, int self, FF x, FF y, FF z,
                                          float &e_lj, float &fx, float &fy, float &fz, float &vir, unsigned &cnt)
{
    float e = 0.f, v = 0.f;
    fx = fy = fz = 0.f;
    // a pair far outside the cutoff is recognised from the leading parts alone (their separation is off by < 1e-5 of L):
    // only the candidates within 1.01 rc pay for the full hi+lo separation
    const float rc2_pre = b.rc2 * 1.0201f + 1e-3f;
#pragma unroll 4
    for (int j = 0; j < N; j++) {
        {
            const float ax = wrapf(x.hi - c.xh[j], b.L, b.invL), ay = wrapf(y.hi - c.yh[j], b.L, b.invL);
            float az = z.hi - c.zh[j];
            if (PZ) az = wrapf(az, b.Lz, b.invLz);
            if (!(fmaf(az, az, fmaf(ay, ay, ax * ax)) < rc2_pre)) continue;
        }
        const float dx = sepwrapf(x.hi, x.lo, c.xh[j], c.xl[j], b.L, b.invL);
        const float dy = sepwrapf(y.hi, y.lo, c.yh[j], c.yl[j], b.L, b.invL);
        const float dz = PZ ? sepwrapf(z.hi, z.lo, c.zh[j], c.zl[j], b.Lz, b.invLz) : sepf(z.hi, z.lo, c.zh[j], c.zl[j]);
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (r2 < b.rc2 && j != self) {
            const float i2 = rcpf(r2);
            const float i6 = i2 * i2 * i2;
            e += fmaf(i6, i6, -i6);
            const float g = i2 * i6 * fmaf(48.f, i6, -24.f);
            fx = fmaf(g, dx, fx); fy = fmaf(g, dy, fy); fz = fmaf(g, dz, fz);
            v += i6 * fmaf(-48.f, i6, 24.f);                 // pressure()'s 24/r^6 - 48/r^12 (SMC.c:712-714)
            cnt++;
        }
    }
    e_lj = 4.f * e;
    vir = v;
}

// one molecule against the surface (SMC.c:729-813, 862-895): energy (already *4), force ADDED, and the two wall virials
__device__ __forceinline__ float wall_f32(const BoxF &b, const float *__restrict__ Wf, FF x, FF y, FF z,
                                          float &fx, float &fy, float &fz, float &vir_ref, float &vir_int)
{
    // distance to the nearer wall: rz + Lz/2 wrapped by Lz, clamped outside the box (SMC.c:735-739)
    const float halfLz_hi = 0.5f * b.Lz.hi, halfLz_lo = 0.5f * b.Lz.lo;
    float dz = wrapf((z.hi + halfLz_hi) + (z.lo + halfLz_lo), b.Lz, b.invLz);
    const float zf = z.hi + z.lo;
    if (zf <= -(halfLz_hi + halfLz_lo)) dz = 0.0001f;
    else if (zf >= halfLz_hi + halfLz_lo) dz = -0.0001f;
    float i2 = rcpf(dz * dz), i6 = i2 * i2 * i2, a6 = b.a0 * i6;
    float e = fmaf(a6, i6, -b.b0 * i6);
    fz = fmaf(i2 * i6 * fmaf(48.f, a6, -24.f * b.b0), dz, fz);
    vir_int = i6 * fmaf(-48.f, a6, 24.f * b.b0);             // flat-wall term once
    // wallsPressure as written: dz from rz + L/2 (sic), no clamp, flat-wall term once per in-cutoff site
    const float dzr = wrapf((z.hi + 0.5f * b.L.hi) + (z.lo + 0.5f * b.L.lo), b.Lz, b.invLz);
    const float j2 = rcpf(dzr * dzr), j6 = j2 * j2 * j2;
    const float zterm_ref = j6 * fmaf(-48.f, b.a0 * j6, 24.f * b.b0);
    vir_ref = 0.f;
    const int MM = b.M * b.M;
    const float *S = Wf + 2 * MM;                             // site coordinates as pairs: [4][MM] = x.hi, x.lo, y.hi, y.lo
    for (int i = 0; i < b.M; i++)
        for (int j = 0; j < b.M; j++) {
            const int m = j + i * b.M;
            const float dx = sepwrapf(x.hi, x.lo, S[m], S[MM + m], b.L, b.invL);
            const float dy = sepwrapf(y.hi, y.lo, S[2 * MM + m], S[3 * MM + m], b.L, b.invL);
            const float ca = Wf[2 * m], cb = Wf[2 * m + 1];
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (r2 < b.rc2) {
                i2 = rcpf(r2); i6 = i2 * i2 * i2; a6 = ca * i6;
                e += fmaf(a6, i6, -cb * i6);
                const float g = i2 * i6 * fmaf(48.f, a6, -24.f * cb);
                fx = fmaf(g, dx, fx); fy = fmaf(g, dy, fy); fz = fmaf(g, dz, fz);
                vir_int += i6 * fmaf(-48.f, a6, 24.f * cb);
            }
            const float r2r = fmaf(dzr, dzr, fmaf(dy, dy, dx * dx));
            if (r2r < b.rc2) {
                const float k2 = rcpf(r2r), k6 = k2 * k2 * k2;
                vir_ref += k6 * fmaf(-48.f, ca * k6, 24.f * cb) + zterm_ref;
            }
        }
    return 4.f * e;
}

__device__ __forceinline__ void stage_f32(const ConfF &c, int j, double X, double Y, double Z)
{
    const FF a = ff_split(X), b = ff_split(Y), d = ff_split(Z);
    c.xh[j] = a.hi; c.xl[j] = a.lo; c.yh[j] = b.hi; c.yl[j] = b.lo; c.zh[j] = d.hi; c.zl[j] = d.lo;
}

// ---- static evaluation ---------------------------------------------------------------------------------------
template <bool PZ>
__device__ __forceinline__ void evaluate_f32_body(const DevChains &d, const EvalOut &o)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    extern __shared__ double sm[];
    double *scratch = sm;                                     // 8*32 doubles
    ConfF c;
    c.carve(reinterpret_cast<float *>(sm + 8 * 32), Npad);
    float *Wf = c.zl + Npad;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b64 = make_box(cp, d.M, 1.0);
    const BoxF b = make_boxf(b64);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    const double *P = d.pos + (size_t)chain * 3 * Npad;
    for (int j = tid; j < N; j += T_) stage_f32(c, j, P[j], P[Npad + j], P[2 * Npad + j]);
    for (int m = tid; m < 2 * d.M * d.M; m += T_) Wf[m] = (float)W[m];
    for (int m = tid; m < d.M * d.M; m += T_) {               // site (i, j) sits at (i L/M, j L/M) in the wall plane (SMC.c:745-750)
        const int MM = d.M * d.M, si = m / d.M, sj = m - si * d.M;
        const FF sx = ff_split(si * (b64.L / d.M)), sy = ff_split(sj * (b64.L / d.M));
        Wf[2 * MM + m] = sx.hi; Wf[3 * MM + m] = sx.lo; Wf[4 * MM + m] = sy.hi; Wf[5 * MM + m] = sy.lo;
    }
    __syncthreads();
    double tot[kTot] = {0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned cnt = 0;
    for (int i = tid; i < N; i += T_) {
        const FF x{c.xh[i], c.xl[i]}, y{c.yh[i], c.yl[i]}, z{c.zh[i], c.zl[i]};
        float e_lj, fx, fy, fz, vir, wx = 0.f, wy = 0.f, wz = 0.f, e_wall = 0.f, vr = 0.f, vi = 0.f;
        pairs_f32<PZ>(b, c, N, i, x, y, z, e_lj, fx, fy, fz, vir, cnt);
        if (b.wall) e_wall = wall_f32(b, Wf, x, y, z, wx, wy, wz, vr, vi);
        const size_t q = (size_t)chain * Npad + i, q3 = (size_t)chain * 3 * Npad + i;
        if (o.e_lj) o.e_lj[q] = e_lj;
        if (o.e_wall) o.e_wall[q] = e_wall;
        if (o.f_lj) { o.f_lj[q3] = fx; o.f_lj[q3 + Npad] = fy; o.f_lj[q3 + 2 * Npad] = fz; }
        if (o.f_wall) { o.f_wall[q3] = wx; o.f_wall[q3 + Npad] = wy; o.f_wall[q3 + 2 * Npad] = wz; }
        tot[0] += 0.5 * (double)e_lj; tot[1] += (double)e_wall; tot[2] += 0.5 * (double)vir;
        tot[3] += (double)vr; tot[4] += (double)vi;
    }
    block_sum<kTot>(tot, scratch);
    if (tid == 0 && o.totals) {
        double *t = o.totals + (size_t)chain * kTot;
        for (int k = 0; k < kTot; k++) t[k] = tot[k];
    }
}

__global__ void k_evaluate_f32(DevChains d, EvalOut o)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) evaluate_f32_body<true>(d, o);
    else evaluate_f32_body<false>(d, o);
}

static __host__ __device__ inline size_t fp32_eval_smem(int Npad, int M) { return (size_t)8 * 32 * sizeof(double) + (size_t)(6 * Npad + 6 * M * M + 4) * sizeof(float); }
static __host__ __device__ inline size_t fp32_step_smem(int Npad, int M) { return (size_t)8 * 32 * sizeof(double) + (size_t)(12 * Npad + 6 * M * M + 4) * sizeof(float); }

// ---- the all-particle Smart-MC step in FP32 ---------------------------------------------------------------------
// One block per chain; current and proposed configuration as hi/lo pairs in shared memory; forces and displacements
// in the float images of a.F / a.Fn / a.dl (the engine's double buffers, reinterpreted: [3][Npad] floats per chain).
template <bool FED, bool PZ>
__device__ __forceinline__ void allparticle_f32_body(const DevChains &d, const StepArgs &a)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    extern __shared__ double sm[];
    double *scratch = sm;
    __shared__ int s_accept;
    ConfF cur, nxt;
    cur.carve(reinterpret_cast<float *>(sm + 8 * 32), Npad);
    nxt.carve(cur.zl + Npad, Npad);
    float *Wf = nxt.zl + Npad;
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b64 = make_box(cp, d.M, d.step_scale);
    const BoxF b = make_boxf(b64);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    float *Fc = reinterpret_cast<float *>(a.F + (size_t)chain * 3 * Npad);
    float *Fn = reinterpret_cast<float *>(a.Fn + (size_t)chain * 3 * Npad);
    float *DL = reinterpret_cast<float *>(a.dl + (size_t)chain * 3 * Npad);
    for (int j = tid; j < N; j += T_) stage_f32(cur, j, P[j], P[Npad + j], P[2 * Npad + j]);
    for (int m = tid; m < 2 * d.M * d.M; m += T_) Wf[m] = (float)W[m];
    for (int m = tid; m < d.M * d.M; m += T_) {               // site (i, j) sits at (i L/M, j L/M) in the wall plane (SMC.c:745-750)
        const int MM = d.M * d.M, si = m / d.M, sj = m - si * d.M;
        const FF sx = ff_split(si * (b64.L / d.M)), sy = ff_split(sj * (b64.L / d.M));
        Wf[2 * MM + m] = sx.hi; Wf[3 * MM + m] = sx.lo; Wf[4 * MM + m] = sy.hi; Wf[5 * MM + m] = sy.lo;
    }
    __syncthreads();

    const float AoT = (float)(b64.A / b64.T), sigma = (float)sqrt(2.0 * b64.A);
    const double invT = 1.0 / b64.T, quarterAoT = 0.25 * b64.A / b64.T;
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    unsigned cnt = 0;
    long long nacc = 0;
    double U;
    {   // U and F of the start configuration (the FP32 mode never trusts FP64 leftovers: always refreshed)
        double t[1] = {0.0};
        for (int i = tid; i < N; i += T_) {
            const FF x{cur.xh[i], cur.xl[i]}, y{cur.yh[i], cur.yl[i]}, z{cur.zh[i], cur.zl[i]};
            float e_lj, fx, fy, fz, vir, ew = 0.f, vr, vi;
            pairs_f32<PZ>(b, cur, N, i, x, y, z, e_lj, fx, fy, fz, vir, cnt);
            if (b.wall) ew = wall_f32(b, Wf, x, y, z, fx, fy, fz, vr, vi);
            Fc[i] = fx; Fc[Npad + i] = fy; Fc[2 * Npad + i] = fz;
            t[0] += 0.5 * (double)e_lj + (double)ew;
        }
        block_sum<1>(t, scratch);
        U = t[0];
        cnt = 0;
    }
    for (int s = 0; s < a.nsteps; s++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)s;
        const size_t sc = (size_t)s * d.C + chain;
        // ---- proposal: d_i = F_i A/T + xi_i ; r' = wrap(r + d) (SMC.c:307-316 for every molecule)
        for (int i = tid; i < N; i += T_) {
            float g0, g1, g2;
            if (FED) {
                const double *xi = a.xi + sc * 3 * N;
                g0 = (float)xi[3 * i]; g1 = (float)xi[3 * i + 1]; g2 = (float)xi[3 * i + 2];
            } else {
                double h0, h1, h2;
                rng_particle_gauss_f32(id, step, (uint32_t)i, h0, h1, h2);      // single-precision values in doubles
                g0 = (float)h0 * sigma; g1 = (float)h1 * sigma; g2 = (float)h2 * sigma;
            }
            const float dX = fmaf(Fc[i], AoT, g0), dY = fmaf(Fc[Npad + i], AoT, g1), dZ = fmaf(Fc[2 * Npad + i], AoT, g2);
            DL[i] = dX; DL[Npad + i] = dY; DL[2 * Npad + i] = dZ;
            FF x = ff_add(FF{cur.xh[i], cur.xl[i]}, dX), y = ff_add(FF{cur.yh[i], cur.yl[i]}, dY), z = ff_add(FF{cur.zh[i], cur.zl[i]}, dZ);
            // wrap into the primary cell: subtract k*L from the pair (k is almost always 0)
            float k = rintf((x.hi + x.lo) * b.invL);
            if (k != 0.f) { x = ff_add(x, -k * b.L.hi); x = ff_add(x, -k * b.L.lo); }
            k = rintf((y.hi + y.lo) * b.invL);
            if (k != 0.f) { y = ff_add(y, -k * b.L.hi); y = ff_add(y, -k * b.L.lo); }
            if (PZ) {
                k = rintf((z.hi + z.lo) * b.invLz);
                if (k != 0.f) { z = ff_add(z, -k * b.Lz.hi); z = ff_add(z, -k * b.Lz.lo); }
            }
            nxt.xh[i] = x.hi; nxt.xl[i] = x.lo; nxt.yh[i] = y.hi; nxt.yl[i] = y.lo; nxt.zh[i] = z.hi; nxt.zl[i] = z.lo;
        }
        __syncthreads();
        // ---- forces and energy at the proposal, Metropolis-Hastings sums (SMC.c:326-329 over all molecules)
        double t[3] = {0.0, 0.0, 0.0};
        float t1 = 0.f, t2 = 0.f;
        for (int i = tid; i < N; i += T_) {
            const FF x{nxt.xh[i], nxt.xl[i]}, y{nxt.yh[i], nxt.yl[i]}, z{nxt.zh[i], nxt.zl[i]};
            float e_lj, fx, fy, fz, vir, ew = 0.f, vr, vi;
            pairs_f32<PZ>(b, nxt, N, i, x, y, z, e_lj, fx, fy, fz, vir, cnt);
            if (b.wall) ew = wall_f32(b, Wf, x, y, z, fx, fy, fz, vr, vi);
            Fn[i] = fx; Fn[Npad + i] = fy; Fn[2 * Npad + i] = fz;
            const float ox = Fc[i], oy = Fc[Npad + i], oz = Fc[2 * Npad + i];
            t[0] += 0.5 * (double)e_lj + (double)ew;
            t1 += fmaf(DL[i], fx + ox, fmaf(DL[Npad + i], fy + oy, DL[2 * Npad + i] * (fz + oz)));
            t2 += (fx - ox) * (fx + ox) + (fy - oy) * (fy + oy) + (fz - oz) * (fz + oz);
        }
        t[1] = (double)t1; t[2] = (double)t2;
        block_sum<3>(t, scratch);
        const double lnap = -((t[0] - U) + 0.5 * t[1] + t[2] * quarterAoT) * invT;
        if (tid == 0) {
            double uu;
            if (FED) uu = a.u[sc];
            else { uint32_t o_; rng_step_scalars(id, step, o_, uu); }
            const int acc = (log(uu) < lnap) ? 1 : 0;           // a NaN exponent (overflowed r^-12) rejects
            s_accept = acc;
            if (a.lnap) a.lnap[sc] = lnap;
            if (a.accepted) a.accepted[sc] = (unsigned char)acc;
        }
        __syncthreads();
        if (s_accept) {
            ConfF tc = cur; cur = nxt; nxt = tc;
            float *tp = Fc; Fc = Fn; Fn = tp;
            U = t[0];
            nacc++;
        }
        __syncthreads();
    }
    for (int j = tid; j < N; j += T_) {
        P[j] = (double)cur.xh[j] + (double)cur.xl[j];
        P[Npad + j] = (double)cur.yh[j] + (double)cur.yl[j];
        P[2 * Npad + j] = (double)cur.zh[j] + (double)cur.zl[j];
    }
    double cc[1] = {(double)cnt};
    block_sum<1>(cc, scratch);
    if (tid == 0) {
        d.E[chain] = U;
        d.nacc[chain] += nacc;
        d.ntri[chain] += a.nsteps;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsteps * (unsigned long long)N * (N - 1));
            atomicAdd(d.pair_counts + 1, (unsigned long long)cc[0]);
            atomicAdd(d.pair_counts + 2, (unsigned long long)(a.nsteps + 1) * (unsigned long long)N * (N - 1));
        }
    }
}

template <bool FED>
__global__ void k_allparticle_f32(DevChains d, StepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) allparticle_f32_body<FED, true>(d, a);
    else allparticle_f32_body<FED, false>(d, a);
}

}  // namespace smcb
