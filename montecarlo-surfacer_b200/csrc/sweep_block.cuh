// sweep_block.cuh — the FAST sweep (oneParticleMoves, SMC.c:278-351) for chains too large for the
// warp-per-chain kernel (N > 512): one thread BLOCK per chain.
//
// Same Markov chain as the reference (same proposal, acceptance expression, visiting order, random inputs) and
// the same arithmetic as k_sweep_cached's general path, organised like the reference's own trial: energy/force
// of the trial particle at its old position, proposal, energy/force at the proposal, Metropolis-Hastings test.
// Each of the two evaluations is the block-wide version of the screened pair loop: the chain's positions live in
// shared memory (exact doubles + box-unit floats), every thread screens its share of the partners in packed
// FP32 (two per instruction), evaluates its hits in FP64 and the four sums are reduced over the block in a fixed
// order.  No per-particle caches here (they would not fit shared memory at N = 4096 next to the positions);
// a trial costs two O(N / threads) passes and a handful of barriers.
#pragma once

namespace smcb {

struct BlockSweepSmem {
    double *x, *y, *z;        // exact positions                       [3][Npad]
    float *fx, *fy, *fz;      // box units, screen precision           [3][Npad]
    double *g0, *g1, *g2, *lu;   // random inputs of the current batch of trials [4][T]
    double *scratch;          // block_sum scratch                     [8*32]
    __device__ __forceinline__ void carve(double *base, int Npad, int T)
    {
        x = base; y = x + Npad; z = y + Npad;
        g0 = z + Npad; g1 = g0 + T; g2 = g1 + T; lu = g2 + T;
        scratch = lu + T;
        fx = reinterpret_cast<float *>(scratch + 8 * 32); fy = fx + Npad; fz = fy + Npad;
    }
    static __host__ __device__ size_t bytes(int Npad, int T)
    {
        return (size_t)(3 * Npad + 4 * T + 8 * 32) * sizeof(double) + (size_t)3 * Npad * sizeof(float);
    }
};

// block-wide sum of four values with ONE barrier: warp totals by the transposed butterfly, one slot per warp in
// `buf` (4 * nwarps doubles), then every thread adds the slots in warp order (fixed order, identical everywhere).
// The caller alternates between two buffers so that a buffer is never rewritten before everyone has read it.
__device__ __forceinline__ void block_sum4_once(double (&v)[4], double *buf)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    warp_sum4(lane, v[0], v[1], v[2], v[3]);
    if (lane == 0) { buf[4 * warp] = v[0]; buf[4 * warp + 1] = v[1]; buf[4 * warp + 2] = v[2]; buf[4 * warp + 3] = v[3]; }
    __syncthreads();
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int w = 0; w < nwarps; w++) { t0 += buf[4 * w]; t1 += buf[4 * w + 1]; t2 += buf[4 * w + 2]; t3 += buf[4 * w + 3]; }
    v[0] = t0; v[1] = t1; v[2] = t2; v[3] = t3;
}

// energy (already *4) and force of a particle at p against all others (index `self` skipped) and the surface;
// every thread returns the block totals.  nin: this thread's partners inside the cutoff (for the pair counter).
template <bool PZ>
__device__ __forceinline__ void block_eval_point(const Box &b, const ScreenConsts &sc, const BlockSweepSmem &s, const double *__restrict__ W,
                                                 int N, int Npad, int self, double px, double py, double pz, int which,
                                                 double &U, double &Fx, double &Fy, double &Fz, unsigned &nin)
{
    const int tid = threadIdx.x, T_ = blockDim.x;
    const float qx = (float)(px * b.invL), qy = (float)(py * b.invL), qz = (float)(pz * b.invL);
    const float2 ax = make_float2(qx, qx), ay = make_float2(qy, qy), az = make_float2(qz, qz);
    const float2 MG = make_float2(12582912.f, 12582912.f);
    const float2 *X2 = reinterpret_cast<const float2 *>(s.fx), *Y2 = reinterpret_cast<const float2 *>(s.fy),
                 *Z2 = reinterpret_cast<const float2 *>(s.fz);
    double v[4] = {0.0, 0.0, 0.0, 0.0};               // e, fx, fy, fz
    // phase 1: screen this thread's partners (pairs j2 = tid, tid + T, ...), two bits per iteration; at most 16
    // iterations (N <= 6016 with 256 threads).  Phase 2 evaluates the hits: kept out of the screen loop so that a
    // warp runs max-over-lanes exact evaluations, not one per iteration in which any lane has a hit.
    unsigned hits = 0;
    const int nit = (Npad / 2 - tid + T_ - 1) / T_;   // this thread's iterations
#pragma unroll 4
    for (int it = 0; it < nit; it++) {
        const int j2 = tid + it * T_;
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
        if (r2.x < sc.rc2s) hits |= 1u << (2 * it);
        if (r2.y < sc.rc2s) hits |= 2u << (2 * it);
    }
    while (hits) {                                     // ascending j within the thread
        const int bit = __ffs(hits) - 1;
        hits &= hits - 1;
        const int j = 2 * (tid + T_ * (bit >> 1)) + (bit & 1);
        double et, gx, gy, gz;
        if (j != self && j < N && pair_exact(b, px, py, pz, s.x[j], s.y[j], s.z[j], et, gx, gy, gz)) {
            v[0] += et; v[1] += gx; v[2] += gy; v[3] += gz;
            nin++;
        }
    }
    __syncwarp();
    double ew = 0.0, fzw = 0.0;
    if (b.wall) {
        const double dzw = wall_dz<false>(b, pz);
        add_zwall(b, dzw, ew, fzw);                    // flat wall: uniform, added after the reduction
        if (dzw * dzw < b.rc2) {                       // sites: thread m owns site m (M*M <= threads for M <= 16)
            const int MM = b.M * b.M;
            const double dw = b.L / b.M;
            for (int m = tid; m < MM; m += T_) {
                const int i = m / b.M, j = m - i * b.M;
                const double dx = min_image<false>(px - i * dw, b.L, b.invL);
                const double dy = min_image<false>(py - j * dw, b.L, b.invL);
                const double r2w = fma(dzw, dzw, fma(dy, dy, dx * dx));
                if (r2w < b.rc2) {
                    const double i2 = fast_rcp(r2w);
                    const double i6 = i2 * i2 * i2;
                    const double a6 = W[2 * m] * i6;
                    v[0] += fma(a6, i6, -W[2 * m + 1] * i6);
                    const double g = i2 * i6 * fma(48.0, a6, -24.0 * W[2 * m + 1]);
                    v[1] = fma(g, dx, v[1]); v[2] = fma(g, dy, v[2]); v[3] = fma(g, dzw, v[3]);
                }
            }
        }
    }
    block_sum4_once(v, s.scratch + 64 * which);       // 4 * 16 warps per buffer; old / proposed position alternate
    U = 4.0 * (v[0] + ew); Fx = v[1]; Fy = v[2]; Fz = v[3] + fzw;
}

template <bool FED, bool PZ>
__device__ __forceinline__ void sweep_block_body(const DevChains &d, const SweepArgs &a)
{
    const int chain = blockIdx.x, N = d.N, Npad = d.Npad, tid = threadIdx.x, T_ = blockDim.x;
    extern __shared__ double sm[];
    BlockSweepSmem s;
    s.carve(sm, Npad, T_);
    const smcb_chain_params &cp = chain_params(d, chain);
    const Box b = make_box(cp, d.M, d.step_scale);
    const ScreenConsts sc = make_screen(b, d.extent ? d.extent + 2 * chain : nullptr);
    const double *W = d.W + (size_t)cp.wall * 2 * d.M * d.M;
    double *P = d.pos + (size_t)chain * 3 * Npad;
    for (int j = tid; j < Npad; j += T_) {
        const bool in = j < N;
        const double X = in ? P[j] : 0.0, Y = in ? P[Npad + j] : 0.0, Z = in ? P[2 * Npad + j] : 0.0;
        s.x[j] = X; s.y[j] = Y; s.z[j] = Z;
        s.fx[j] = (float)(X * b.invL); s.fy[j] = (float)(Y * b.invL); s.fz[j] = in ? (float)(Z * b.invL) : 3.0e18f;
    }
    __syncthreads();

    const double AoT = b.A / b.T, sigma = sqrt(2.0 * b.A), quarterAoT = 0.25 * AoT, invT = 1.0 / b.T;
    const RngId id{a.rng.k0, a.rng.k1, a.rng.chain0 + (uint32_t)chain};
    double E = d.E[chain];
    int nacc = 0;
    unsigned cnt = 0;

    for (int sw = 0; sw < a.nsweeps; sw++) {
        const unsigned long long step = a.rng.step0 + (unsigned long long)sw;
        const size_t sci = (size_t)sw * d.C + chain;
        const int nacc0 = nacc;
        long long offset;                              // int offset = rand();  SMC.c:290
        if (FED) offset = a.offset[sci];
        else { uint32_t o; double unused; rng_step_scalars(id, step, o, unused); offset = o; }
        const int off = (int)(offset % N);
        for (int nn0 = 0; nn0 < N; nn0 += T_) {
            // random inputs of the next T trials, one per thread (trial nn visits particle (nn + off) % N, SMC.c:294)
            {
                const int nn = nn0 + tid;
                if (nn < N) {
                    int n = nn + off;
                    if (n >= N) n -= N;
                    double g0, g1, g2, ul;
                    if (FED) {
                        const double *dsp = a.displ + sci * 3 * N;
                        g0 = dsp[3 * n]; g1 = dsp[3 * n + 1]; g2 = dsp[3 * n + 2];
                        ul = a.u[sci * N + nn];
                    } else {
                        rng_particle_gauss_f32(id, step, (uint32_t)n, g0, g1, g2);
                        g0 *= sigma; g1 *= sigma; g2 *= sigma;
                        ul = rng_particle_uniform(id, step, (uint32_t)n);
                    }
                    s.g0[tid] = g0; s.g1[tid] = g1; s.g2[tid] = g2; s.lu[tid] = log(ul);
                }
            }
            __syncthreads();
            const int tmax = min(T_, N - nn0);
            for (int t = 0; t < tmax; t++) {
                int n = nn0 + t + off;
                if (n >= N) n -= N;
                const double px = s.x[n], py = s.y[n], pz = s.z[n];
                double Um, Fmx, Fmy, Fmz, Un, Fnx, Fny, Fnz;
                block_eval_point<PZ>(b, sc, s, W, N, Npad, n, px, py, pz, 0, Um, Fmx, Fmy, Fmz, cnt);          // SMC.c:300-304
                const double dX = fma(Fmx, AoT, s.g0[t]), dY = fma(Fmy, AoT, s.g1[t]), dZ = fma(Fmz, AoT, s.g2[t]);   // SMC.c:307-309
                const double qx = min_image<false>(px + dX, b.L, b.invL), qy = min_image<false>(py + dY, b.L, b.invL);   // SMC.c:311-316
                double qz = pz + dZ;
                if (PZ) qz = min_image<false>(qz, b.Lz, b.invLz);
                block_eval_point<PZ>(b, sc, s, W, N, Npad, n, qx, qy, qz, 1, Un, Fnx, Fny, Fnz, cnt);          // SMC.c:319-321
                // SMC.c:326-335: accept iff u < exp(-(Un-Um + d.(Fn+Fm)/2 + (Fn^2-Fm^2) A/(4T))/T)
                const double f2 = fma(Fnx, Fnx, fma(Fny, Fny, Fnz * Fnz)) - fma(Fmx, Fmx, fma(Fmy, Fmy, Fmz * Fmz));
                const double dr = fma(dX, Fnx + Fmx, fma(dY, Fny + Fmy, dZ * (Fnz + Fmz)));
                const double xarg = -((Un - Um) + 0.5 * dr + f2 * quarterAoT) * invT;
                const bool acc = (s.lu[t] < xarg) && (xarg > -745.1332191019411);
                if (acc) {
                    if (tid == 0) {
                        s.x[n] = qx; s.y[n] = qy; s.z[n] = qz;
                        s.fx[n] = (float)(qx * b.invL); s.fy[n] = (float)(qy * b.invL); s.fz[n] = (float)(qz * b.invL);
                    }
                    E += Un - Um;                       // SMC.c:341
                    nacc++;
                }
                if (FED && a.accepted != nullptr && tid == 0) a.accepted[sci * N + nn0 + t] = acc ? 1 : 0;
                __syncthreads();                        // the move is visible before the next trial reads positions
            }
        }
        if (a.trace_E != nullptr && tid == 0) { a.trace_E[sci] = E; a.trace_acc[sci] = nacc - nacc0; }
    }

    for (int j = tid; j < N; j += T_) { P[j] = s.x[j]; P[Npad + j] = s.y[j]; P[2 * Npad + j] = s.z[j]; }
    double c[1] = {(double)cnt};
    block_sum<1>(c, s.scratch);
    if (tid == 0) {
        d.E[chain] = E;
        d.nacc[chain] += nacc;
        d.ntri[chain] += (long long)a.nsweeps * N;
        if (d.pair_counts) {
            atomicAdd(d.pair_counts, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));
            atomicAdd(d.pair_counts + 1, (unsigned long long)c[0]);
            atomicAdd(d.pair_counts + 2, (unsigned long long)a.nsweeps * 2ull * N * (N - 1));    // no caches: old + proposed position
        }
    }
}

template <bool FED>
__global__ void __launch_bounds__(512) k_sweep_block(DevChains d, SweepArgs a)
{
    if (chain_params(d, blockIdx.x).flags & SMCB_PERIODIC_Z) sweep_block_body<FED, true>(d, a);
    else sweep_block_body<FED, false>(d, a);
}

}  // namespace smcb
