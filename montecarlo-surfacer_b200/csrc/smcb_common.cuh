// smcb_common.cuh — device-side physics shared by every kernel of libsmcb200.
//
// Each helper is templated on STRICT:
//   STRICT = true   the reference's IEEE operation order, written exactly as the
//                   C expressions of /root/reference/SMC.c read (true divisions,
//                   left-to-right products).  Translation units that instantiate
//                   STRICT code are compiled with --fmad=false, so no FMA is ever
//                   contracted and per-particle results are bit-identical to the
//                   reference built with -ffp-contract=off.
//   STRICT = false  the fused formulation: multiply by 1/L, one reciprocal per
//                   pair, FMAs everywhere (17 + 16 algorithmic flops per pair,
//                   SURVEY.md §8d).  Agrees with the reference to ~1e-15 relative.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/smcb200.h"

namespace smcb {

constexpr unsigned FULL = 0xffffffffu;

// per-chain constants, expanded once per kernel from smcb_chain_params
struct Box {
    double L, Lz, invL, invLz, rc2, a0, b0, T, A;
    bool wall, pz;
    int M;
};

__device__ __forceinline__ Box make_box(const smcb_chain_params &p, int M, double step_scale)
{
    Box b;
    b.L = p.L; b.Lz = p.Lz; b.invL = 1.0 / p.L; b.invLz = 1.0 / p.Lz;
    b.rc2 = p.rc2; b.a0 = p.zwall_a; b.b0 = p.zwall_b; b.T = p.T; b.A = p.A * step_scale;
    b.wall = (p.flags & SMCB_WALL) != 0; b.pz = (p.flags & SMCB_PERIODIC_Z) != 0;
    b.M = M;
    return b;
}

// d - P*rint(d/P)   (SMC.c:568 and every other minimum-image line)
template <bool STRICT>
__device__ __forceinline__ double min_image(double d, double P, double invP)
{
    if (STRICT) {
        return d - P * rint(d / P);
    } else {
        // k is a small integer, so P*k is exact and the FMA rounds once, like the
        // reference; d*invP can differ from d/P only within an ulp of a half-integer,
        // where |d - P*k| ~ P/2 is outside the cutoff either way.
        // rint by the 1.5*2^52 trick: two DADDs on the FP64 pipe (16 cycles) instead of FRND.F64 on the XU
        // pipe (37 cycles latency, 9 cycles per warp instruction; profiles/r01/microbench_fp64_lat.txt).
        // Exact round-to-nearest-even for |d/P| < 2^51, i.e. the same k as rint().
        const double k = __dsub_rn(__dadd_rn(d * invP, 6755399441055744.0), 6755399441055744.0);
        return fma(-P, k, d);
    }
}

// separation a-b under the box rules; returns r^2 = dx*dx + dy*dy + dz*dz
template <bool STRICT>
__device__ __forceinline__ double pair_sep(const Box &b, double ax, double ay, double az,
                                           double bx, double by, double bz,
                                           double &dx, double &dy, double &dz)
{
    dx = min_image<STRICT>(ax - bx, b.L, b.invL);
    dy = min_image<STRICT>(ay - by, b.L, b.invL);
    dz = az - bz;                                   // slab: z is never wrapped (SMC.c:571-572)
    if (b.pz) dz = min_image<STRICT>(dz, b.Lz, b.invLz);
    if (STRICT) return dx * dx + dy * dy + dz * dz;
    return fma(dz, dz, fma(dy, dy, dx * dx));
}

// One in-cutoff 12-6 pair with coefficients (ca, cb): energy term e (to be
// multiplied by 4 by the caller, like `return V*4`, SMC.c:582) and g = -(dV/dr)/r.
// ca = cb = 1 for molecule-molecule pairs (SMC.c:577-578, 612-613), (W[2m],
// W[2m+1]) for surface sites (SMC.c:756-757, 805-806).
template <bool STRICT, bool UNIT>
__device__ __forceinline__ void lj_terms(double r2, double ca, double cb, double &e, double &g)
{
    if (STRICT) {
        double r6 = r2 * r2 * r2;
        double r8 = r2 * r2 * r2 * r2;
        if (UNIT) {
            e = 1.0 / (r6 * r6) - 1.0 / r6;
            g = 48.0 / (r8 * r2 * r2 * r2) - 24.0 / r8;
        } else {
            e = ca / (r6 * r6) - cb / r6;
            g = 48.0 * ca / (r8 * r2 * r2 * r2) - 24.0 * cb / r8;
        }
    } else {
        double i2 = 1.0 / r2;
        double i6 = i2 * i2 * i2;
        if (UNIT) {
            e = fma(i6, i6, -i6);
            g = i2 * i6 * fma(48.0, i6, -24.0);
        } else {
            double a6 = ca * i6;
            e = fma(a6, i6, -cb * i6);
            g = i2 * i6 * fma(48.0, a6, -24.0 * cb);
        }
    }
}

// pressure()'s pair term 24/r^6 - 48/r^12 (SMC.c:712-714)
template <bool STRICT>
__device__ __forceinline__ double virial_term(double r2)
{
    if (STRICT) {
        double r6 = r2 * r2 * r2;
        return 24.0 / r6 - 48.0 / (r6 * r6);
    } else {
        double i2 = 1.0 / r2;
        double i6 = i2 * i2 * i2;
        return i6 * fma(-48.0, i6, 24.0);
    }
}

// signed distance to the nearer wall with the reference's clamp (SMC.c:735-739)
template <bool STRICT>
__device__ __forceinline__ double wall_dz(const Box &b, double rz)
{
    double dz = rz + b.Lz / 2;
    dz = min_image<STRICT>(dz, b.Lz, b.invLz);
    if (rz <= -b.Lz / 2.0) dz = 0.0001;
    else if (rz >= b.Lz / 2) dz = -0.0001;
    return dz;
}

// flat z-wall: e = a0/dz^12 - b0/dz^6 (no cutoff), g*dz is the z force (SMC.c:740-741, 787-789)
template <bool STRICT>
__device__ __forceinline__ void zwall_terms(const Box &b, double dz, double &e, double &g)
{
    if (STRICT) {
        double z6 = dz * dz * dz * dz * dz * dz;
        e = b.a0 / (z6 * z6) - b.b0 / z6;
        double z8 = dz * dz * dz * dz * dz * dz * dz * dz;
        g = 48.0 * b.a0 / (z8 * dz * dz * dz * dz * dz * dz) - 24.0 * b.b0 / z8;
    } else {
        double i2 = 1.0 / (dz * dz);
        double i6 = i2 * i2 * i2;
        double a6 = b.a0 * i6;
        e = fma(a6, i6, -b.b0 * i6);
        g = i2 * i6 * fma(48.0, a6, -24.0 * b.b0);
    }
}

// One particle against the whole surface, thread-serial, in the reference's
// order: flat wall first, then sites m = i*M + j ascending (SMC.c:729-763,
// 773-813).  Returns the energy WITHOUT the final *4 and ADDS the force.
template <bool STRICT>
__device__ __forceinline__ double wall_particle(const Box &b, const double *__restrict__ W,
                                                double rx, double ry, double rz,
                                                double &fx, double &fy, double &fz)
{
    double acc = 0.0, e, g;
    const double dw = b.L / b.M;
    const double dz = wall_dz<STRICT>(b, rz);
    zwall_terms<STRICT>(b, dz, e, g);
    acc += e;
    fz += g * dz;
    for (int i = 0; i < b.M; i++)
        for (int j = 0; j < b.M; j++) {
            const int m = j + i * b.M;
            double dx = min_image<STRICT>(rx - i * dw, b.L, b.invL);
            double dy = min_image<STRICT>(ry - j * dw, b.L, b.invL);
            double r2 = STRICT ? (dx * dx + dy * dy + dz * dz) : fma(dz, dz, fma(dy, dy, dx * dx));
            if (r2 < b.rc2) {
                lj_terms<STRICT, false>(r2, W[2 * m], W[2 * m + 1], e, g);
                acc += e;
                fx += g * dx;
                fy += g * dy;
                fz += g * dz;
            }
        }
    return acc;
}

// wallsPressure()'s per-particle sum AS THE REFERENCE WRITES IT (SMC.c:862-895):
// dz = rz + L/2 (sic), wrapped by Lz, no clamp; the flat-wall term is added once
// per in-cutoff site.
template <bool STRICT>
__device__ __forceinline__ double wall_virial_ref(const Box &b, const double *__restrict__ W,
                                                  double rx, double ry, double rz)
{
    double acc = 0.0;
    const double dw = b.L / b.M;
    double dz = rz + b.L / 2;
    dz = min_image<STRICT>(dz, b.Lz, b.invLz);
    for (int i = 0; i < b.M; i++)
        for (int j = 0; j < b.M; j++) {
            const int m = j + i * b.M;
            double dx = min_image<STRICT>(rx - i * dw, b.L, b.invL);
            double dy = min_image<STRICT>(ry - j * dw, b.L, b.invL);
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < b.rc2) {
                double r6 = r2 * r2 * r2;
                acc += 24.0 * W[2 * m + 1] / r6 - 48.0 * W[2 * m] / (r6 * r6);
                double z6 = dz * dz * dz * dz * dz * dz;
                acc += 24.0 * b.b0 / z6 - 48.0 * b.a0 / (z6 * z6);
            }
        }
    return acc;
}

// The wall virial the reference MEANT to compute: sum of r dV/dr over the surface terms of one
// particle with the same geometry as wallsEnergySingle / wallsForce - distance to the nearer wall from rz + Lz/2
// with the clamp (SMC.c:735-739), the flat-wall term ONCE, and every site inside the cutoff.
template <bool STRICT>
__device__ __forceinline__ double wall_virial_intended(const Box &b, const double *__restrict__ W, double rx, double ry, double rz)
{
    const double dw = b.L / b.M;
    const double dz = wall_dz<STRICT>(b, rz);
    const double z6 = dz * dz * dz * dz * dz * dz;
    double acc = 24.0 * b.b0 / z6 - 48.0 * b.a0 / (z6 * z6);
    for (int i = 0; i < b.M; i++)
        for (int j = 0; j < b.M; j++) {
            const int m = j + i * b.M;
            const double dx = min_image<STRICT>(rx - i * dw, b.L, b.invL);
            const double dy = min_image<STRICT>(ry - j * dw, b.L, b.invL);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < b.rc2) {
                const double r6 = r2 * r2 * r2;
                acc += 24.0 * W[2 * m + 1] / r6 - 48.0 * W[2 * m] / (r6 * r6);
            }
        }
    return acc;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// block-wide sum of up to 8 doubles at once; result valid in every thread.
// `scratch` needs 8 * 32 doubles of shared memory.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] = warp_sum(v[q]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) scratch[q * 32 + warp] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double t = (lane < nwarps) ? scratch[q * 32 + lane] : 0.0;
        v[q] = warp_sum(t);
    }
}

}  // namespace smcb
