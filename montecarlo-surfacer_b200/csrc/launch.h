// launch.h — host-side launch entry points of the two kernel translation units.
#pragma once
#include "kernels.cuh"

namespace smcb {

// kernels_fast.cu (FMA contraction on) and kernels_strict.cu (--fmad=false)
// each define one set; `strict` picks the set at run time in engine.cu.
#define SMCB_DECLARE_LAUNCHERS(SUFFIX)                                                              \
    cudaError_t launch_sweep_##SUFFIX(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st); \
    cudaError_t launch_allparticle_##SUFFIX(bool fed, const DevChains &d, const StepArgs &a, cudaStream_t st);

SMCB_DECLARE_LAUNCHERS(fast)
SMCB_DECLARE_LAUNCHERS(strict)
cudaError_t launch_evaluate_strict(const DevChains &d, const EvalOut &o, cudaStream_t st);   // FAST: launch_evaluate_fast_screened

int evaluate_fast_parts(const DevChains &d);
cudaError_t launch_evaluate_fast_screened(const DevChains &d, const EvalOut &o, int parts, double *partials, unsigned *tickets, cudaStream_t st);
cudaError_t launch_evaluate_f32(const DevChains &d, const EvalOut &o, cudaStream_t st);
cudaError_t launch_allparticle_f32(bool fed, const DevChains &d, const StepArgs &a, cudaStream_t st);
cudaError_t launch_gather(const DevChains &d, const GatherArgs &g, cudaStream_t st);
cudaError_t launch_gather_chains(const DevChains &d, const GatherArgs &g, cudaStream_t st);
cudaError_t launch_gather_moments(const DevChains &d, const GatherArgs &g, cudaStream_t st);
cudaError_t launch_chain_extent(const DevChains &d, float *extent, int *flag, cudaStream_t st);
cudaError_t launch_adapt_step(smcb_chain_params *params, long long *nacc, long long *ntri, int C, double target, double gain,
                              double a_min, double a_max, cudaStream_t st);
cudaError_t launch_energy_from_totals(const double *totals, double *E, int C, cudaStream_t st);
cudaError_t launch_aos_to_soa(const double *aos, double *soa, int C, int N, int Npad, int ncomp, cudaStream_t st);
cudaError_t launch_soa_to_aos(const double *soa, double *aos, int C, int N, int Npad, int ncomp, cudaStream_t st);
cudaError_t launch_dfma_peak(double *out, int blocks, int threads, int iters, cudaStream_t st);

// largest N the warp-per-chain sweep kernel is instantiated for
constexpr int kSweepMaxN = 512;
// largest N of the block-per-chain FAST sweep (positions twice in shared memory next to 10 KB of scratch)
constexpr int kSweepBlockMaxN = 6016;

}  // namespace smcb
