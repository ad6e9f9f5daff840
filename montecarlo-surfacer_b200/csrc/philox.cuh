// philox.cuh — counter-based random streams of the engine (Philox4x32-10).
//
// Replaces the reference's libc rand() (SMC.c:290,335; matematicose.c:188-189):
// a chain's numbers are a pure function of (seed, global chain id, step,
// particle), so any sharding of chains over GPUs reproduces the same job, and
// the CPU restatement (oracle/smc_oracle.c: orc_rng_particle,
// orc_rng_step_scalars) replays the identical stream.
//   key     = (seed_lo, seed_hi)
//   counter = (step_lo, step_hi, chain, particle | tag << 28)
//   tag 0,1 : the two blocks behind a particle's three N(0,1) numbers (STRICT kernels: Box-Muller in double
//             precision on 53-bit uniforms, replayed exactly by the oracle); the FAST kernels take all three
//             from block 0 with a SINGLE-precision Box-Muller (32-bit uniforms, logf/sincospif, ~1e-7 relative;
//             the tail reaches 6.6 sigma): the transform ran on the scarce FP64 pipe and cost as much as the
//             half-shell pair screen of the all-particle step
//   tag 2   : the particle's trial uniform (sweep kernel)
//   tag 3   : per-step scalars (sweep offset, whole-chain uniform); particle = 0
#pragma once
#include <cstdint>

namespace smcb {

struct RngId {
    uint32_t k0, k1;       // seed
    uint32_t chain;        // global chain id
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform strictly inside (0,1)
__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi)
{
    const unsigned long long v = ((static_cast<unsigned long long>(hi) << 32) | lo) >> 11;
    return (static_cast<double>(v) + 0.5) * (1.0 / 9007199254740992.0);
}

// three standard normals (Box-Muller on 53-bit uniforms) for (step, particle)
__device__ __forceinline__ void rng_particle_gauss(const RngId &id, unsigned long long step, uint32_t particle,
                                                   double &g0, double &g1, double &g2)
{
    uint32_t a[4], b[4];
    philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), id.chain, particle, id.k0, id.k1, a);
    philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), id.chain, particle | (1u << 28), id.k0, id.k1, b);
    const double u1 = u53(a[0], a[1]), u2 = u53(a[2], a[3]);
    const double u3 = u53(b[0], b[1]), u4 = u53(b[2], b[3]);
    const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    g0 = r1 * c;
    g1 = r1 * s;
    g2 = r2 * cospi(2.0 * u4);
}

// FAST variant: three standard normals from ONE Philox block, single-precision transform, returned as doubles
__device__ __forceinline__ void rng_particle_gauss_f32(const RngId &id, unsigned long long step, uint32_t particle,
                                                       double &g0, double &g1, double &g2)
{
    uint32_t a[4];
    philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), id.chain, particle, id.k0, id.k1, a);
    // (a + 1/2) / 2^32 in (0, 1]: one rounding in the fma, small values keep their full precision (the tail)
    const float u1 = fmaf(static_cast<float>(a[0]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float u3 = fmaf(static_cast<float>(a[2]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u3));
    float sn, cs;
    sincospif(static_cast<float>(a[1]) * 4.6566128730773926e-10f, &sn, &cs);      // angle 2 pi a1 / 2^32
    g0 = static_cast<double>(r1 * cs);
    g1 = static_cast<double>(r1 * sn);
    g2 = static_cast<double>(r2 * cospif(static_cast<float>(a[3]) * 4.6566128730773926e-10f));
}

__device__ __forceinline__ double rng_particle_uniform(const RngId &id, unsigned long long step, uint32_t particle)
{
    uint32_t c[4];
    philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), id.chain, particle | (2u << 28), id.k0, id.k1, c);
    return u53(c[0], c[1]);
}

__device__ __forceinline__ void rng_step_scalars(const RngId &id, unsigned long long step, uint32_t &offset, double &u)
{
    uint32_t c[4];
    philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), id.chain, 3u << 28, id.k0, id.k1, c);
    offset = c[2] & 0x7fffffffu;
    u = u53(c[0], c[1]);
}

}  // namespace smcb
