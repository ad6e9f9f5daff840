// launchers.inl — included by kernels_fast.cu and kernels_strict.cu with
// SMCB_TU_STRICT = false / true and SMCB_TU_SUFFIX = fast / strict.
namespace smcb {

#define SMCB_CAT2(a, b) a##b
#define SMCB_CAT(a, b) SMCB_CAT2(a, b)

static inline int eval_threads(int N)
{
    int t = ((N + 31) / 32) * 32;
    return t > 512 ? 512 : (t < 64 ? 64 : t);     // 100-120 registers per thread: 512 is the largest block that launches
}

#if !SMCB_TU_IS_STRICT
// blocks per chain of the FAST evaluation: enough to put at least ~2 blocks on every SM
int evaluate_fast_parts(const DevChains &d)
{
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int parts = 1;
    while (parts < 16 && d.C * parts < 2 * sms && d.N / (parts * 2) >= 128) parts *= 2;
    return parts;
}

cudaError_t launch_evaluate_fast_screened(const DevChains &d, const EvalOut &o, int parts, double *partials, unsigned *tickets, cudaStream_t st)
{
    const size_t smem = StepSmem::bytes(d.Npad);
    cudaError_t err = cudaFuncSetAttribute(k_evaluate_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int per = (d.N + parts - 1) / parts;
    int threads = ((per + 31) / 32) * 32;
    threads = threads > 512 ? 512 : (threads < 64 ? 64 : threads);
    EvalFastArgs ea{parts, partials, tickets};
    k_evaluate_fast<<<d.C * parts, threads, smem, st>>>(d, o, ea);
    return cudaGetLastError();
}
#endif

#if SMCB_TU_IS_STRICT
cudaError_t SMCB_CAT(launch_evaluate_, SMCB_TU_SUFFIX)(const DevChains &d, const EvalOut &o, cudaStream_t st)
{
    const size_t smem = (size_t)(3 * d.Npad + 8 * 32) * sizeof(double);
    auto kern = k_evaluate<SMCB_TU_STRICT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<d.C, eval_threads(d.N), smem, st>>>(d, o);
    return cudaGetLastError();
}
#endif

template <int K>
static cudaError_t sweep_k(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st)
{
#if SMCB_TU_IS_STRICT
    const size_t smem = (size_t)3 * d.Npad * sizeof(double);
    if (fed) k_sweep<K, true, true><<<d.C, 32, smem, st>>>(d, a);
    else     k_sweep<K, true, false><<<d.C, 32, smem, st>>>(d, a);
#else
    const int MMpad = (d.M * d.M + 3) & ~3;
    // Two FAST kernels: k_sweep_spec (segment-speculative, sweep_spec.cuh) for gas-like states, k_sweep_cached for condensed
    // ones (every accepted move voids a quarter of a segment's speculation there).  The engine passes what the previous
    // launch saw; SMCB_SWEEP_KERNEL = cached | spec pins one of them (A/B measurements).
    static const int pin = [] { const char *e = getenv("SMCB_SWEEP_KERNEL"); return !e ? 0 : (!strcmp(e, "cached") ? 1 : (!strcmp(e, "spec") ? 2 : 0)); }();
    const bool use_spec = pin ? pin == 2 : a.dense_hint != 1;
    size_t smem = use_spec ? SpecSmem<K>::bytes(MMpad) : ChainSmem::bytes(32 * K, MMpad);
    if (const char *env = getenv("SMCB_SWEEP_SMEM_PAD")) smem += (size_t)atoi(env);     // occupancy experiments (profiles/)
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t err;
    auto go = [&](auto kern) -> cudaError_t {
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return err;
        kern<<<d.C, 32, smem, st>>>(d, a);
        return cudaSuccess;
    };
    if (use_spec) err = fed ? go(k_sweep_spec<K, true>) : go(k_sweep_spec<K, false>);
    else          err = fed ? go(k_sweep_cached<K, true>) : go(k_sweep_cached<K, false>);
    if (err != cudaSuccess) return err;
#endif
    return cudaGetLastError();
}

#if !SMCB_TU_IS_STRICT
// ---- SMCB_FP32 (fp32_mode.cuh) ----
cudaError_t launch_evaluate_f32(const DevChains &d, const EvalOut &o, cudaStream_t st)
{
    const size_t smem = fp32_eval_smem(d.Npad, d.M);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t err = cudaFuncSetAttribute(k_evaluate_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k_evaluate_f32<<<d.C, eval_threads(d.N), smem, st>>>(d, o);
    return cudaGetLastError();
}

cudaError_t launch_allparticle_f32(bool fed, const DevChains &d, const StepArgs &a, cudaStream_t st)
{
    const size_t smem = fp32_step_smem(d.Npad, d.M);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t err;
    int threads = ((d.N + 31) / 32) * 32;
    threads = threads > 256 ? 256 : (threads < 64 ? 64 : threads);
    if (fed) {
        if ((err = cudaFuncSetAttribute(k_allparticle_f32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_allparticle_f32<true><<<d.C, threads, smem, st>>>(d, a);
    } else {
        if ((err = cudaFuncSetAttribute(k_allparticle_f32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_allparticle_f32<false><<<d.C, threads, smem, st>>>(d, a);
    }
    return cudaGetLastError();
}

// N > 512: one block per chain (sweep_block.cuh)
static cudaError_t sweep_block_launch(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st)
{
    // default: the batch-speculative kernel, one warp per trial (sweep_block_spec.cuh); SMCB_BLOCK_SWEEP=serial keeps
    // the trial-by-trial block kernel below reachable for comparison
    const char *which = getenv("SMCB_BLOCK_SWEEP");
    if (!(which && strcmp(which, "serial") == 0)) {
        // one trial per warp; SMCB_BLOCK_SPEC_TPW=2: two (512 threads, every molecule pair loaded once for both points; slower)
        int tpw = 1;
        if (const char *env = getenv("SMCB_BLOCK_SPEC_TPW")) { const int v = atoi(env); if (v == 1 || v == 2) tpw = v; }
        int threads = 1024 / tpw;
        if (const char *env = getenv("SMCB_BLOCK_SWEEP_THREADS")) { const int v = atoi(env); if (v >= 64 && v <= 1024 / tpw && v % 32 == 0) threads = v; }
        const size_t smem = BlockSpecSmem::bytes(d.Npad);
        if (smem > 227 * 1024 || BlockSpecSmem::nf(d.Npad) / 64 > 16 * kBlockSpecWords) return cudaErrorInvalidValue;
        cudaError_t err;
#define SMCB_LAUNCH_BLOCK_SPEC(FEDV, TPWV)                                                                                                     \
        do {                                                                                                                                   \
            if ((err = cudaFuncSetAttribute(k_sweep_block_spec<FEDV, TPWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
            k_sweep_block_spec<FEDV, TPWV><<<d.C, threads, smem, st>>>(d, a);                                                                  \
        } while (0)
        if (fed) { if (tpw == 2) SMCB_LAUNCH_BLOCK_SPEC(true, 2); else SMCB_LAUNCH_BLOCK_SPEC(true, 1); }
        else     { if (tpw == 2) SMCB_LAUNCH_BLOCK_SPEC(false, 2); else SMCB_LAUNCH_BLOCK_SPEC(false, 1); }
#undef SMCB_LAUNCH_BLOCK_SPEC
        return cudaGetLastError();
    }
    int threads = 256;
    if (const char *env = getenv("SMCB_BLOCK_SWEEP_THREADS")) { const int v = atoi(env); if (v >= 64 && v <= 512 && v % 32 == 0) threads = v; }
    // a thread keeps two hit bits per screen iteration in one 32-bit word: at most 16 iterations (block_eval_point)
    while ((d.Npad / 2 + threads - 1) / threads > 16 && threads < 512) threads += 32;
    if ((d.Npad / 2 + threads - 1) / threads > 16) return cudaErrorInvalidValue;
    const size_t smem = BlockSweepSmem::bytes(d.Npad, threads);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t err;
    if (fed) {
        if ((err = cudaFuncSetAttribute(k_sweep_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_sweep_block<true><<<d.C, threads, smem, st>>>(d, a);
    } else {
        if ((err = cudaFuncSetAttribute(k_sweep_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        k_sweep_block<false><<<d.C, threads, smem, st>>>(d, a);
    }
    return cudaGetLastError();
}
#endif

cudaError_t SMCB_CAT(launch_sweep_, SMCB_TU_SUFFIX)(bool fed, const DevChains &d, const SweepArgs &a, cudaStream_t st)
{
#if !SMCB_TU_IS_STRICT
    if (d.N > kSweepMaxN) return sweep_block_launch(fed, d, a, st);
#else
    if (d.N > kSweepMaxN) {                      // the bit-exact sweep beyond one warp's registers: one block per chain
        const char *which = getenv("SMCB_BLOCK_SWEEP");
        if (!(which && strcmp(which, "serial") == 0)) {            // batch-speculative, one warp per trial (same bits)
            const size_t smem = BlockSpecSmem::bytes(d.Npad);
            if (smem > 227 * 1024 || BlockSpecSmem::nf(d.Npad) / 64 > 16 * kBlockSpecWords) return cudaErrorInvalidValue;
            int threads = 512;
            if (const char *env = getenv("SMCB_BLOCK_SWEEP_THREADS")) { const int v = atoi(env); if (v >= 64 && v <= 512 && v % 32 == 0) threads = v; }
            cudaError_t err;
            if (fed) {
                if ((err = cudaFuncSetAttribute(k_sweep_block_strict_spec<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
                k_sweep_block_strict_spec<true><<<d.C, threads, smem, st>>>(d, a);
            } else {
                if ((err = cudaFuncSetAttribute(k_sweep_block_strict_spec<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
                k_sweep_block_strict_spec<false><<<d.C, threads, smem, st>>>(d, a);
            }
            return cudaGetLastError();
        }
        const size_t smem = StrictBlockSmem::bytes(d.Npad);
        if (smem > 227 * 1024) return cudaErrorInvalidValue;
        cudaError_t err;
        if (fed) {
            if ((err = cudaFuncSetAttribute(k_sweep_block_strict<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
            k_sweep_block_strict<true><<<d.C, 256, smem, st>>>(d, a);
        } else {
            if ((err = cudaFuncSetAttribute(k_sweep_block_strict<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
            k_sweep_block_strict<false><<<d.C, 256, smem, st>>>(d, a);
        }
        return cudaGetLastError();
    }
#endif
    const int k = d.Npad / 32;
    if (k <= 1) return sweep_k<1>(fed, d, a, st);
    if (k <= 2) return sweep_k<2>(fed, d, a, st);
    if (k <= 4) return sweep_k<4>(fed, d, a, st);
    if (k <= 8) return sweep_k<8>(fed, d, a, st);
    if (k <= 16) return sweep_k<16>(fed, d, a, st);
    return cudaErrorInvalidValue;
}

#if !SMCB_TU_IS_STRICT
// blocks per chain of the FAST all-particle kernel: 1, or a thread-block cluster when a batch of few
// large chains would leave SMs idle (config 5: 32 chains x N = 4096 per GPU -> clusters of 4)
static int allparticle_cluster(const DevChains &d)
{
    if (const char *env = getenv("SMCB_CLUSTER")) {
        const int v = atoi(env);
        if (v == 1 || v == 2 || v == 4 || v == 8) return v;
    }
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int cl = 1;
    while (cl < 8 && d.C * cl * 2 <= sms && d.N / (cl * 2) >= 256) cl *= 2;
    return cl;
}

template <bool FED, int CL, bool HALF = false>
static cudaError_t allparticle_fast_k(const DevChains &d, const StepArgs &a, cudaStream_t st)
{
    const size_t smem = StepSmem::bytes(d.Npad, d.N, HALF);
    auto kern = k_allparticle_fast<FED, CL, HALF>;
    cudaError_t err;
    if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    // two molecules per thread: blocks half as wide, twice as many resident per SM, so one block's barriers
    // (five per step) overlap with the other blocks' pair loops (measured: 11.4 -> 9.9 ms per 40 steps at N=256)
    int per = (d.N + CL - 1) / CL;
    int threads = ((per / 2 + 31) / 32) * 32;
    threads = threads > 512 ? 512 : (threads < 64 ? 64 : threads);
    if (const char *env = getenv("SMCB_ALLP_THREADS")) { const int v = atoi(env); if (v >= 32 && v <= 512 && v % 32 == 0) threads = v; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)d.C * CL);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, d, a);
}

cudaError_t launch_allparticle_fast(bool fed, const DevChains &d, const StepArgs &a, cudaStream_t st)
{
    switch (allparticle_cluster(d)) {
    case 8: return fed ? allparticle_fast_k<true, 8>(d, a, st) : allparticle_fast_k<false, 8>(d, a, st);
    case 4: return fed ? allparticle_fast_k<true, 4>(d, a, st) : allparticle_fast_k<false, 4>(d, a, st);
    case 2: return fed ? allparticle_fast_k<true, 2>(d, a, st) : allparticle_fast_k<false, 2>(d, a, st);
    default:
        // one block per chain: screen every unordered pair once (half shell) while the hit words fit comfortably
        if (d.N <= 512 && !getenv("SMCB_FULL_SHELL"))
            return fed ? allparticle_fast_k<true, 1, true>(d, a, st) : allparticle_fast_k<false, 1, true>(d, a, st);
        return fed ? allparticle_fast_k<true, 1>(d, a, st) : allparticle_fast_k<false, 1>(d, a, st);
    }
}
#else
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
#endif

}  // namespace smcb
