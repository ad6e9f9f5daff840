// launchers.inl — included by kernels_fast.cu and kernels_strict.cu with
// SMCB_TU_STRICT = false / true and SMCB_TU_SUFFIX = fast / strict.
namespace smcb {

#define SMCB_CAT2(a, b) a##b
#define SMCB_CAT(a, b) SMCB_CAT2(a, b)

static inline int eval_threads(int N)
{
    int t = ((N + 31) / 32) * 32;
    return t > 1024 ? 1024 : (t < 64 ? 64 : t);
}

cudaError_t SMCB_CAT(launch_evaluate_, SMCB_TU_SUFFIX)(const DevChains &d, const EvalOut &o, cudaStream_t st)
{
    const size_t smem = (size_t)(3 * d.Npad + 8 * 32) * sizeof(double);
    auto kern = k_evaluate<SMCB_TU_STRICT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<d.C, eval_threads(d.N), smem, st>>>(d, o);
    return cudaGetLastError();
}

template <int K>
static cudaError_t sweep_k(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st)
{
#if SMCB_TU_IS_STRICT
    const size_t smem = (size_t)3 * d.Npad * sizeof(double);
    if (fed) k_sweep<K, true, true><<<d.C, 32, smem, st>>>(d, a);
    else     k_sweep<K, true, false><<<d.C, 32, smem, st>>>(d, a);
#else
    const int MMpad = (d.M * d.M + 3) & ~3;
    const size_t smem = ChainSmem::bytes(d.Npad, MMpad);
    cudaError_t err;
    if (fed) {
        auto kern = k_sweep_cached<K, true>;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return err;
        kern<<<d.C, 32, smem, st>>>(d, a);
    } else {
        auto kern = k_sweep_cached<K, false>;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return err;
        kern<<<d.C, 32, smem, st>>>(d, a);
    }
#endif
    return cudaGetLastError();
}

cudaError_t SMCB_CAT(launch_sweep_, SMCB_TU_SUFFIX)(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st)
{
    const int k = d.Npad / 32;
    if (k <= 1) return sweep_k<1>(fed, d, a, st);
    if (k <= 2) return sweep_k<2>(fed, d, a, st);
    if (k <= 4) return sweep_k<4>(fed, d, a, st);
    if (k <= 8) return sweep_k<8>(fed, d, a, st);
    if (k <= 16) return sweep_k<16>(fed, d, a, st);
    return cudaErrorInvalidValue;
}

cudaError_t SMCB_CAT(launch_allparticle_, SMCB_TU_SUFFIX)(bool fed, const DevChains &d, const StepArgs &a, cudaStream_t st)
{
    const size_t smem = (size_t)(6 * d.Npad + 8 * 32) * sizeof(double);
    cudaError_t err;
    if (fed) {
        auto kern = k_allparticle<SMCB_TU_STRICT, true>;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        kern<<<d.C, eval_threads(d.N), smem, st>>>(d, a);
    } else {
        auto kern = k_allparticle<SMCB_TU_STRICT, false>;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        kern<<<d.C, eval_threads(d.N), smem, st>>>(d, a);
    }
    return cudaGetLastError();
}

}  // namespace smcb
