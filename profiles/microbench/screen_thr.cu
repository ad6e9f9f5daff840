// screen_thr.cu — cost of the sweep kernel's phase-1 screen (K=8 register-resident slots per lane
// against one broadcast point, minimum image in x,y, cutoff test) in three arithmetic variants:
//   0  FP64, as k_sweep_cached does it (x wrap on the XU pipe via FRND.F64, y wrap by the 2^52 trick)
//   1  FP32 scalar (magic-number rounding, FADD/FMUL/FFMA)
//   2  FP32 packed pairs (add/mul/fma .f32x2: two slots per instruction, sm_100+)
//   3  FP64 with both wraps by the 2^52 trick (no XU)
// Reported: cycles per screen (8 slots) per warp, for W warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o screen_thr screen_thr.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 2048
constexpr int K = 8;

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

template <int V>
__global__ void screen(unsigned *out, long long *cyc, double seed, double rc2s)
{
    const int lane = threadIdx.x & 31;
    double xs[K], ys[K], zs[K];
    float fx[K], fy[K], fz[K];
    float2 px[K / 2], py[K / 2], pz[K / 2];
    for (int k = 0; k < K; k++) {
        xs[k] = -0.5 + (lane * 8 + k) / 256.0; ys[k] = 0.3 - (lane * 5 + k * 3) / 400.0; zs[k] = (k - 4) * 0.7 + lane * 0.01;
        fx[k] = (float)xs[k]; fy[k] = (float)ys[k]; fz[k] = (float)zs[k];
    }
    for (int k = 0; k < K / 2; k++) {
        px[k] = make_float2(fx[2 * k], fx[2 * k + 1]); py[k] = make_float2(fy[2 * k], fy[2 * k + 1]); pz[k] = make_float2(fz[2 * k], fz[2 * k + 1]);
    }
    double qx = seed, qy = -seed, qz = seed * 0.5;
    unsigned acc = 0;
    const float rc2f = (float)rc2s;
    const float MAGIC = 12582912.f;
    long long t0 = clock64();
    for (int it = 0; it < N_IT; it++) {
        unsigned hits = 0;
        if (V == 0 || V == 3) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                double sx = qx - xs[k];
                if (V == 0) sx -= rint(sx);
                else sx -= __dsub_rn(__dadd_rn(sx, 6755399441055744.0), 6755399441055744.0);
                double sy = qy - ys[k];
                sy -= __dsub_rn(__dadd_rn(sy, 6755399441055744.0), 6755399441055744.0);
                const double sz = qz - zs[k];
                if (fma(sz, sz, fma(sy, sy, sx * sx)) < rc2s) hits |= 1u << k;
            }
        } else if (V == 1) {
            const float ax = (float)qx, ay = (float)qy, az = (float)qz;
#pragma unroll
            for (int k = 0; k < K; k++) {
                float sx = ax - fx[k];
                sx -= __fsub_rn(__fadd_rn(sx, MAGIC), MAGIC);
                float sy = ay - fy[k];
                sy -= __fsub_rn(__fadd_rn(sy, MAGIC), MAGIC);
                const float sz = az - fz[k];
                if (fmaf(sz, sz, fmaf(sy, sy, sx * sx)) < rc2f) hits |= 1u << k;
            }
        } else {
            const float ax = (float)qx, ay = (float)qy, az = (float)qz;
            const float2 a2x = make_float2(ax, ax), a2y = make_float2(ay, ay), a2z = make_float2(az, az);
            const float2 M2 = make_float2(MAGIC, MAGIC);
#pragma unroll
            for (int k = 0; k < K / 2; k++) {
                float2 sx = sub2(a2x, px[k]);
                sx = sub2(sx, sub2(add2(sx, M2), M2));
                float2 sy = sub2(a2y, py[k]);
                sy = sub2(sy, sub2(add2(sy, M2), M2));
                const float2 sz = sub2(a2z, pz[k]);
                const float2 r2 = fma2(sz, sz, fma2(sy, sy, mul2(sx, sx)));
                if (r2.x < rc2f) hits |= 1u << (2 * k);
                if (r2.y < rc2f) hits |= 2u << (2 * k);
            }
        }
        const bool any = __any_sync(0xffffffffu, hits != 0);
        acc += hits;
        // next point depends on this trial's outcome (as in the kernel: accept -> positions change)
        qx = any ? qx * 0.999 : qx + 0.001; qy += 0.0007; qz -= 0.0003;
        if (qx > 0.5) qx -= 1.0;
        if (qy > 0.5) qy -= 1.0;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    unsigned *out; long long *cyc, h[8];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const char *names[] = {"FP64 (XU + magic)", "FP32 scalar", "FP32 packed f32x2", "FP64 (magic only)"};
#define RUN(V, W) screen<V><<<1, 32 * 4 * W>>>(out, cyc, 0.123, 9.0 / 1089.0); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-20s warps/SMSP=%d : %7.1f cycles per 8-slot screen per warp, %6.1f per SMSP-screen\n", names[V], W, (double)h[0] / N_IT, (double)h[0] / N_IT / W);
    RUN(0, 1) RUN(0, 2) RUN(0, 3) RUN(0, 4)
    RUN(3, 1) RUN(3, 2) RUN(3, 3) RUN(3, 4)
    RUN(1, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4)
    RUN(2, 1) RUN(2, 2) RUN(2, 3) RUN(2, 4)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
