// fp64_lat.cu — latency / throughput of the instructions the sweep kernel's critical path is made
// of, measured on the device with clock64() (one warp for latency; many warps for throughput).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP>
__global__ void lat(double *out, long long *cyc, double a, double b)
{
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; i++) {
        if (OP == 0) x = fma(x, b, a);                       // DFMA
        if (OP == 1) x = x + b;                              // DADD
        if (OP == 2) x = rint(x * b) + a;                    // DMUL + FRND + DADD
        if (OP == 3) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y + a; }  // MUFU.RCP64H + DADD
        if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;   // 2 SHFL + DADD
        if (OP == 5) x = x * b;                              // DMUL
        if (OP == 6) x = (x < a) ? x + b : x - b;            // DSETP + select path
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// throughput: ILP independent chains per thread
template <int OP, int ILP>
__global__ void thr(double *out, long long *cyc, double a, double b)
{
    double x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = a + threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (OP == 0) x[j] = fma(x[j], b, a);
            if (OP == 2) x[j] = rint(x[j]);
            if (OP == 4) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const char *names[] = {"DFMA", "DADD", "DMUL+FRND.F64+DADD", "MUFU.RCP64H+DADD", "SHFL(x2)+DADD", "DMUL", "DSETP+sel+DADD"};
#define RUN_LAT(OP) lat<OP><<<1, 32>>>(out, cyc, 1.000001, 0.999999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("latency  %-22s %7.2f cycles/iter (1 warp, dependent chain)\n", names[OP], (double)h[0] / N_IT);
    RUN_LAT(0) RUN_LAT(1) RUN_LAT(2) RUN_LAT(3) RUN_LAT(4) RUN_LAT(5) RUN_LAT(6)
#define RUN_THR(OP, ILP, W, LABEL) thr<OP, ILP><<<1, 32 * W>>>(out, cyc, 1.000001, 0.999999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("thruput  %-10s ILP=%d warps/SM=%2d : %6.2f cycles per warp-instr per SMSP\n", LABEL, ILP, W, (double)h[0] / N_IT / ILP / ((W + 3) / 4));
    RUN_THR(0, 8, 4, "DFMA") RUN_THR(0, 8, 8, "DFMA") RUN_THR(0, 8, 16, "DFMA")
    RUN_THR(2, 8, 4, "FRND.F64") RUN_THR(2, 8, 8, "FRND.F64") RUN_THR(2, 8, 16, "FRND.F64")
    RUN_THR(4, 8, 4, "SHFLx2") RUN_THR(4, 8, 16, "SHFLx2")
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
