// kernels_fast.cu — FAST instantiations (FMA contraction on) plus the kernels
// that have no strict/fast distinction (gather, layout transposes, DFMA peak).
#include <cstdlib>
#include <cstring>
#define SMCB_MISC_KERNELS
#include "launch.h"
#define SMCB_TU_IS_STRICT 0
#define SMCB_TU_STRICT false
#define SMCB_TU_SUFFIX fast
#include "launchers.inl"

namespace smcb {

cudaError_t launch_gather(const DevChains &d, const GatherArgs &g, cudaStream_t st)
{
    k_gather<<<d.C, 128, 0, st>>>(d, g);
    k_gather_moments<<<g.ngroups, 256, 0, st>>>(d, g);
    return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(const double *aos, double *soa, int C, int N, int Npad, int ncomp, cudaStream_t st)
{
    const size_t total = (size_t)C * Npad;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    k_aos_to_soa<<<blocks ? blocks : 1, 256, 0, st>>>(aos, soa, C, N, Npad, ncomp);
    return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(const double *soa, double *aos, int C, int N, int Npad, int ncomp, cudaStream_t st)
{
    const size_t total = (size_t)C * N;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    k_soa_to_aos<<<blocks ? blocks : 1, 256, 0, st>>>(soa, aos, C, N, Npad, ncomp);
    return cudaGetLastError();
}

cudaError_t launch_energy_from_totals(const double *totals, double *E, int C, cudaStream_t st)
{
    k_energy_from_totals<<<(C + 255) / 256, 256, 0, st>>>(totals, E, C);
    return cudaGetLastError();
}

cudaError_t launch_gather_chains(const DevChains &d, const GatherArgs &g, cudaStream_t st)      // per-chain part only
{
    k_gather<<<d.C, 128, 0, st>>>(d, g);
    return cudaGetLastError();
}

cudaError_t launch_gather_moments(const DevChains &d, const GatherArgs &g, cudaStream_t st)     // per-group sums, all chains
{
    k_gather_moments<<<g.ngroups, 256, 0, st>>>(d, g);
    return cudaGetLastError();
}

cudaError_t launch_chain_extent(const DevChains &d, float *extent, int *flag, cudaStream_t st)
{
    k_chain_extent<<<(d.C + 7) / 8, 256, 0, st>>>(d.pos, d.params, d.nparams, d.C, d.N, d.Npad, extent, flag);
    return cudaGetLastError();
}

cudaError_t launch_adapt_step(smcb_chain_params *params, long long *nacc, long long *ntri, int C, double target, double gain,
                              double a_min, double a_max, cudaStream_t st)
{
    k_adapt_step<<<(C + 255) / 256, 256, 0, st>>>(params, nacc, ntri, C, target, gain, a_min, a_max);
    return cudaGetLastError();
}

cudaError_t launch_dfma_peak(double *out, int blocks, int threads, int iters, cudaStream_t st)
{
    k_dfma_peak<<<blocks, threads, 0, st>>>(out, iters, 1.0);
    return cudaGetLastError();
}

}  // namespace smcb
