// engine.cu — the C ABI of include/smcb200.h over the kernels in kernels.cuh.
//
// One smcb_engine owns one CUDA stream and every device buffer of a chain batch.
// Host buffers cross the boundary in the reference's layouts (AoS double[3N] per
// chain, SMC.h:84); the AoS<->SoA transposes run on the device.  There is no CPU
// fallback anywhere in this file: every entry point either launches the sm_100a
// kernels or returns an error.
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "launch.h"

using namespace smcb;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SMCB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count)
    {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct smcb_engine {
    int device = 0, C = 0, N = 0, Npad = 0, M = 0, nwalls = 0, ngroups = 1, nparams = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // smcb_sweep_host: the batch is cut into kParts chain blocks, each on its own stream, so that one block's
    // PCIe copies overlap the other blocks' kernels
    static constexpr int kParts = 4;
    cudaStream_t pstream[kParts] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pev[kParts] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pstart = nullptr;
    unsigned long long *pairs_pinned = nullptr;     // pinned landing area of the pair counters
    uint64_t params_hash = 0;                       // FNV-1a of the chain parameters and wall tables (checkpoint guard)
    bool obs_reduced = false;                       // the block holds an in-place all-reduced total: no further gathers until a reset
    DevBuf<smcb_chain_params> params;
    DevBuf<double> W, pos, E, stage, F, Fn, dl, e_lj, f_lj, e_wall, f_wall, totals, moments, peak_out;
    DevBuf<double> fed_a, fed_b;            // host-fed random inputs
    DevBuf<double> cache_out;               // test hook (smcb_debug_capture_cache)
    bool capture_cache = false;
    bool wall_virial_intended = false;      // which wall virial the gathered pressure uses (smcb_obs_set_wall_virial)
    bool totals_valid = false;              // totals hold the last smcb_evaluate / smcb_gather
    DevBuf<long long> nacc, ntri, fed_off;
    DevBuf<unsigned long long> pairs, counters;
    DevBuf<unsigned char> fed_acc;
    DevBuf<int> rbin, trace_acc, extent_flag;
    DevBuf<float> extent;                   // [C][2] box-unit extent of every chain at upload (FP32 screen bound)
    int *flag_pinned = nullptr;
    DevBuf<double> eval_partials, chain_mom;
    DevBuf<unsigned> eval_tickets;
    DevBuf<double> trace_E;
    uint64_t seed = 0x5eed5eedull, step = 0;
    uint32_t chain0 = 0;
    double step_scale = 1.0;
    bool have_params = false, have_pos = false, energy_valid = false, forces_valid = false;
    int nebins = 64;
    double e_lo = -8.0, e_hi = 2.0;
    float last_ms = 0.f;
    int last_launches = 0;
    unsigned long long last_pairs[16] = {0};
    int sweep_dense = 0;                    // the last FAST sweep launch found > 2 % of the pairs inside the cutoff

    DevChains chains()
    {
        DevChains d;
        d.C = C; d.N = N; d.Npad = Npad; d.M = M;
        d.params = params.p; d.nparams = nparams; d.W = W.p; d.pos = pos.p; d.E = E.p;
        d.nacc = nacc.p; d.ntri = ntri.p; d.step_scale = step_scale; d.pair_counts = pairs.p;
        d.extent = extent.p;
        return d;
    }
    size_t u64_per_group() const { return (size_t)2 * SMCB_NCX * SMCB_NCX * SMCB_NCZ + SMCB_NCZ + nebins + 1; }
    size_t f64_per_group() const { return 5; }
};

static int check(smcb_engine *e)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    cudaError_t err = cudaSetDevice(e->device);
    if (err != cudaSuccess) return fail(SMCB_ERR_CUDA, "cudaSetDevice(%d): %s", e->device, cudaGetErrorString(err));
    return SMCB_OK;
}

static int need_ready(smcb_engine *e)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->have_params) return fail(SMCB_ERR_STATE, "smcb_set_params has not been called");
    if (!e->have_pos) return fail(SMCB_ERR_STATE, "smcb_set_positions has not been called");
    return SMCB_OK;
}

// After new positions reach the device: every chain's extent in box units (the FAST kernels' FP32 screen derives its
// error bound from it) and a check that the coordinates can be screened in single precision at all.
static int positions_uploaded(smcb_engine *e)
{
    if (!e->have_params) return SMCB_OK;                 // extents need L: computed when the parameters arrive
    CK(launch_chain_extent(e->chains(), e->extent.p, e->extent_flag.p, e->stream));
    CK(cudaMemcpyAsync(e->flag_pinned, e->extent_flag.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (*e->flag_pinned) {
        CK(cudaMemsetAsync(e->extent_flag.p, 0, sizeof(int), e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->have_pos = false;
        return fail(SMCB_ERR_ARG, "positions contain NaN or coordinates more than 2^20 box lengths from the origin");
    }
    return SMCB_OK;
}

namespace {
struct CkptHeader {
    char magic[8];                 // "SMCB200\0"
    uint32_t version, C, N, M, ngroups, nebins;
    uint32_t chain0, pad_;            // pad_: the sweep kernel hint (sweep_dense)
    uint64_t seed, step;
    double step_scale, e_lo, e_hi;
    uint64_t n_counters, n_moments;
    uint64_t params_hash;          // version 2: FNV-1a of the chain parameters and wall tables the run was made with
};

template <typename T>
int put(FILE *f, const T *dev, size_t n, cudaStream_t st, std::vector<unsigned char> &buf)
{
    buf.resize(n * sizeof(T));
    if (cudaMemcpyAsync(buf.data(), dev, buf.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    return fwrite(buf.data(), 1, buf.size(), f) == buf.size() ? 0 : -2;
}

}  // namespace

extern "C" {

const char *smcb_last_error(void) { return g_err.c_str(); }

int smcb_create(smcb_engine **out, int device, int nchains, int N, int M)
{
    if (!out) return fail(SMCB_ERR_ARG, "out is null");
    *out = nullptr;
    if (nchains <= 0 || N <= 1 || M <= 0) return fail(SMCB_ERR_ARG, "need nchains>0, N>1, M>0 (got %d, %d, %d)", nchains, N, M);
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(SMCB_ERR_NODEVICE, "no CUDA device (%s); libsmcb200 has no CPU path", err == cudaSuccess ? "count=0" : cudaGetErrorString(err));
    if (device < 0 || device >= ndev) return fail(SMCB_ERR_ARG, "device %d out of range (%d present)", device, ndev);
    CK(cudaSetDevice(device));
    smcb_engine *e = new (std::nothrow) smcb_engine();
    if (!e) return fail(SMCB_ERR_ARG, "out of host memory");
    e->device = device; e->C = nchains; e->N = N; e->Npad = ((N + 31) / 32) * 32; e->M = M;
    const size_t cn = (size_t)nchains * e->Npad;
    cudaError_t a = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (a == cudaSuccess) a = cudaEventCreate(&e->ev0);
    if (a == cudaSuccess) a = cudaEventCreate(&e->ev1);
    if (a == cudaSuccess) a = cudaEventCreateWithFlags(&e->pstart, cudaEventDisableTiming);
    for (int p = 0; p < smcb_engine::kParts && a == cudaSuccess; p++) {
        a = cudaStreamCreateWithFlags(&e->pstream[p], cudaStreamNonBlocking);
        if (a == cudaSuccess) a = cudaEventCreateWithFlags(&e->pev[p], cudaEventDisableTiming);
    }
    if (a == cudaSuccess) a = cudaMallocHost(&e->pairs_pinned, 16 * sizeof(unsigned long long));
    if (a == cudaSuccess) a = cudaMallocHost(&e->flag_pinned, sizeof(int));
    if (a == cudaSuccess) a = e->extent.ensure((size_t)2 * nchains);
    if (a == cudaSuccess) a = e->extent_flag.ensure(1);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->extent_flag.p, 0, sizeof(int), e->stream);
    if (a == cudaSuccess) a = e->pos.ensure(3 * cn);
    if (a == cudaSuccess) a = e->E.ensure(nchains);
    if (a == cudaSuccess) a = e->nacc.ensure(nchains);
    if (a == cudaSuccess) a = e->ntri.ensure(nchains);
    if (a == cudaSuccess) a = e->pairs.ensure(16);
    if (a == cudaSuccess) a = e->totals.ensure((size_t)kTot * nchains);
    if (a == cudaSuccess) a = e->rbin.ensure((size_t)nchains * N);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->E.p, 0, nchains * sizeof(double), e->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->nacc.p, 0, nchains * sizeof(long long), e->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->ntri.p, 0, nchains * sizeof(long long), e->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->pairs.p, 0, 16 * sizeof(unsigned long long), e->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(e->rbin.p, 0, (size_t)nchains * N * sizeof(int), e->stream);
    if (a == cudaSuccess) a = cudaStreamSynchronize(e->stream);
    if (a != cudaSuccess) {
        int rc = fail(SMCB_ERR_CUDA, "smcb_create: %s", cudaGetErrorString(a));
        smcb_destroy(e);
        return rc;
    }
    *out = e;
    return SMCB_OK;
}

int smcb_destroy(smcb_engine *e)
{
    if (!e) return SMCB_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->params.release(); e->W.release(); e->pos.release(); e->E.release(); e->stage.release();
    e->F.release(); e->Fn.release(); e->dl.release(); e->e_lj.release(); e->f_lj.release();
    e->e_wall.release(); e->f_wall.release(); e->totals.release(); e->moments.release();
    e->peak_out.release(); e->fed_a.release(); e->fed_b.release(); e->nacc.release(); e->ntri.release();
    e->cache_out.release(); e->chain_mom.release(); e->eval_partials.release(); e->eval_tickets.release(); e->trace_E.release(); e->trace_acc.release(); e->fed_off.release(); e->pairs.release(); e->counters.release(); e->fed_acc.release(); e->rbin.release();
    for (int p = 0; p < smcb_engine::kParts; p++) {
        if (e->pstream[p]) { cudaStreamSynchronize(e->pstream[p]); cudaStreamDestroy(e->pstream[p]); }
        if (e->pev[p]) cudaEventDestroy(e->pev[p]);
    }
    if (e->pstart) cudaEventDestroy(e->pstart);
    if (e->pairs_pinned) cudaFreeHost(e->pairs_pinned);
    if (e->flag_pinned) cudaFreeHost(e->flag_pinned);
    e->extent.release(); e->extent_flag.release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return SMCB_OK;
}

int smcb_device_info(smcb_engine *e, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes)
{
    int rc = check(e);
    if (rc) return rc;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, e->device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
    return SMCB_OK;
}

// Clear the ACCUMULATORS (counters and moments).  Rbin - the voxel every particle was in at the last gather,
// localDensityAndMobility's `Rbin`, SMC.c:921-924 - is chain state, not an accumulator: it survives a reset, so the
// mobility counts of the next accumulation window are those of an uninterrupted run.  with_rbin clears it too (a new
// parameter set / a new observable configuration: the reference callocs it at the start of sMC, SMC.c:54).
static int obs_alloc(smcb_engine *e, bool with_rbin)
{
    CK(e->counters.ensure(e->u64_per_group() * e->ngroups));
    CK(e->moments.ensure(e->f64_per_group() * e->ngroups));
    CK(cudaMemsetAsync(e->counters.p, 0, e->u64_per_group() * e->ngroups * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(e->moments.p, 0, e->f64_per_group() * e->ngroups * sizeof(double), e->stream));
    if (with_rbin) CK(cudaMemsetAsync(e->rbin.p, 0, (size_t)e->C * e->N * sizeof(int), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->obs_reduced = false;
    return SMCB_OK;
}

static uint64_t fnv1a(const void *data, size_t n, uint64_t h = 0xcbf29ce484222325ull)
{
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

int smcb_set_params(smcb_engine *e, const smcb_chain_params *p, int nparams, const double *W, int nwalls, int ngroups)
{
    int rc = check(e);
    if (rc) return rc;
    if (!p || (nparams != 1 && nparams != e->C)) return fail(SMCB_ERR_ARG, "nparams must be 1 or nchains (%d), got %d", e->C, nparams);
    if (ngroups <= 0) return fail(SMCB_ERR_ARG, "ngroups must be positive");
    bool any_wall = false;
    for (int i = 0; i < nparams; i++) {
        if (!(p[i].L > 0) || !(p[i].Lz > 0) || !(p[i].T > 0) || !(p[i].A > 0) || !(p[i].rc2 > 0))
            return fail(SMCB_ERR_ARG, "chain %d: L, Lz, T, A, rc2 must be positive", i);
        if (p[i].flags & SMCB_WALL) {
            any_wall = true;
            if ((int)p[i].wall >= nwalls) return fail(SMCB_ERR_ARG, "chain %d: wall table %u >= nwalls %d", i, p[i].wall, nwalls);
        }
        if ((int)p[i].group >= ngroups) return fail(SMCB_ERR_ARG, "chain %d: group %u >= ngroups %d", i, p[i].group, ngroups);
    }
    if (any_wall && (!W || nwalls <= 0)) return fail(SMCB_ERR_ARG, "SMCB_WALL chains need W and nwalls>0");
    CK(e->params.ensure(nparams));
    CK(cudaMemcpyAsync(e->params.p, p, nparams * sizeof(*p), cudaMemcpyHostToDevice, e->stream));
    const size_t wlen = (size_t)(nwalls > 0 ? nwalls : 1) * 2 * e->M * e->M;
    CK(e->W.ensure(wlen));
    CK(cudaMemsetAsync(e->W.p, 0, wlen * sizeof(double), e->stream));
    if (W && nwalls > 0) CK(cudaMemcpyAsync(e->W.p, W, wlen * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->nparams = nparams; e->nwalls = nwalls; e->ngroups = ngroups;
    e->have_params = true; e->energy_valid = false; e->forces_valid = false;
    e->params_hash = fnv1a(p, (size_t)nparams * sizeof(*p));
    if (W && nwalls > 0) e->params_hash = fnv1a(W, wlen * sizeof(double), e->params_hash);
    if ((rc = obs_alloc(e, true))) return rc;
    return e->have_pos ? positions_uploaded(e) : SMCB_OK;      // L may have changed: extents are in box units
}

int smcb_set_positions(smcb_engine *e, const double *R)
{
    int rc = check(e);
    if (rc) return rc;
    if (!R) return fail(SMCB_ERR_ARG, "R is null");
    const size_t n = (size_t)e->C * 3 * e->N;
    CK(e->stage.ensure(n));
    CK(cudaMemcpyAsync(e->stage.p, R, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(launch_aos_to_soa(e->stage.p, e->pos.p, e->C, e->N, e->Npad, 3, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->have_pos = true; e->energy_valid = false; e->forces_valid = false;
    return positions_uploaded(e);
}

int smcb_broadcast_positions(smcb_engine *e, const double *R0)
{
    int rc = check(e);
    if (rc) return rc;
    if (!R0) return fail(SMCB_ERR_ARG, "R0 is null");
    const size_t one = (size_t)3 * e->N;
    CK(e->stage.ensure(one * e->C));
    CK(cudaMemcpyAsync(e->stage.p, R0, one * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    // doubling copies on the device
    for (size_t have = 1; have < (size_t)e->C;) {
        const size_t take = (have * 2 <= (size_t)e->C) ? have : (size_t)e->C - have;
        CK(cudaMemcpyAsync(e->stage.p + have * one, e->stage.p, take * one * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
        have += take;
    }
    CK(launch_aos_to_soa(e->stage.p, e->pos.p, e->C, e->N, e->Npad, 3, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->have_pos = true; e->energy_valid = false; e->forces_valid = false;
    return positions_uploaded(e);
}

int smcb_get_positions(smcb_engine *e, double *R)
{
    int rc = check(e);
    if (rc) return rc;
    if (!R) return fail(SMCB_ERR_ARG, "R is null");
    if (!e->have_pos) return fail(SMCB_ERR_STATE, "no positions set");
    const size_t n = (size_t)e->C * 3 * e->N;
    CK(e->stage.ensure(n));
    CK(launch_soa_to_aos(e->pos.p, e->stage.p, e->C, e->N, e->Npad, 3, e->stream));
    CK(cudaMemcpyAsync(R, e->stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

int smcb_set_rng(smcb_engine *e, uint64_t seed, uint32_t chain0, uint64_t step0)
{
    int rc = check(e);
    if (rc) return rc;
    e->seed = seed; e->chain0 = chain0; e->step = step0;
    return SMCB_OK;
}

int smcb_set_step_scale(smcb_engine *e, double scale)
{
    int rc = check(e);
    if (rc) return rc;
    if (!(scale > 0)) return fail(SMCB_ERR_ARG, "scale must be positive");
    e->step_scale = scale;
    return SMCB_OK;
}

// --------------------------------------------------------------- evaluation
static int run_evaluate(smcb_engine *e, int mode, const EvalOut &o)
{
    DevChains d = e->chains();
    d.step_scale = 1.0;
    d.pair_counts = nullptr;
    if (mode == SMCB_STRICT) {
        CK(launch_evaluate_strict(d, o, e->stream));
        return SMCB_OK;
    }
    if (mode == SMCB_FP32) {
        CK(launch_evaluate_f32(d, o, e->stream));
        return SMCB_OK;
    }
    // FAST: packed-FP32 screened pair loop, several blocks per chain when the batch is small
    const int parts = evaluate_fast_parts(d);
    CK(e->eval_partials.ensure((size_t)e->C * parts * kTot));
    if (e->eval_tickets.n < (size_t)e->C) {
        CK(e->eval_tickets.ensure(e->C));
        CK(cudaMemsetAsync(e->eval_tickets.p, 0, e->C * sizeof(unsigned), e->stream));
    }
    CK(launch_evaluate_fast_screened(d, o, parts, e->eval_partials.p, e->eval_tickets.p, e->stream));
    return SMCB_OK;
}

static int fetch(smcb_engine *e, const double *soa, double *host, int ncomp)
{
    const size_t n = (size_t)e->C * e->N * ncomp;
    CK(e->stage.ensure((size_t)e->C * 3 * e->N));
    CK(launch_soa_to_aos(soa, e->stage.p, e->C, e->N, e->Npad, ncomp, e->stream));
    CK(cudaMemcpyAsync(host, e->stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

int smcb_evaluate(smcb_engine *e, int mode, double *e_lj, double *f_lj, double *e_wall, double *f_wall,
                  double *U_lj, double *U_wall, double *vir_lj, double *vir_wall_ref)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (mode != SMCB_FAST && mode != SMCB_STRICT && mode != SMCB_FP32) return fail(SMCB_ERR_ARG, "bad mode %d", mode);
    const size_t cn = (size_t)e->C * e->Npad;
    EvalOut o{};
    if (e_lj) { CK(e->e_lj.ensure(cn)); o.e_lj = e->e_lj.p; }
    if (f_lj) { CK(e->f_lj.ensure(3 * cn)); o.f_lj = e->f_lj.p; }
    if (e_wall) { CK(e->e_wall.ensure(cn)); o.e_wall = e->e_wall.p; }
    if (f_wall) { CK(e->f_wall.ensure(3 * cn)); o.f_wall = e->f_wall.p; }
    o.totals = e->totals.p;
    CK(cudaEventRecord(e->ev0, e->stream));
    if ((rc = run_evaluate(e, mode, o))) return rc;
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    e->last_launches = 1;
    e->totals_valid = true;
    if (e_lj && (rc = fetch(e, o.e_lj, e_lj, 1))) return rc;
    if (f_lj && (rc = fetch(e, o.f_lj, f_lj, 3))) return rc;
    if (e_wall && (rc = fetch(e, o.e_wall, e_wall, 1))) return rc;
    if (f_wall && (rc = fetch(e, o.f_wall, f_wall, 3))) return rc;
    if (U_lj || U_wall || vir_lj || vir_wall_ref) {
        std::vector<double> t((size_t)kTot * e->C);
        CK(cudaMemcpyAsync(t.data(), e->totals.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        for (int c = 0; c < e->C; c++) {
            if (U_lj) U_lj[c] = t[kTot * c];
            if (U_wall) U_wall[c] = t[kTot * c + 1];
            if (vir_lj) vir_lj[c] = t[kTot * c + 2];
            if (vir_wall_ref) vir_wall_ref[c] = t[kTot * c + 3];
        }
    }
    return SMCB_OK;
}

// E <- energy(R) + wallsEnergy(R)   (SMC.c:48)
static int refresh_energy(smcb_engine *e, int mode)
{
    EvalOut o{};
    o.totals = e->totals.p;
    int rc = run_evaluate(e, mode, o);
    if (rc) return rc;
    CK(launch_energy_from_totals(e->totals.p, e->E.p, e->C, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->energy_valid = true;
    return SMCB_OK;
}

int smcb_refresh_energy(smcb_engine *e, int mode)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (mode != SMCB_FAST && mode != SMCB_STRICT && mode != SMCB_FP32) return fail(SMCB_ERR_ARG, "bad mode %d", mode);
    return refresh_energy(e, mode);
}

int smcb_get_chain_state(smcb_engine *e, double *E, int64_t *naccept, int64_t *ntrials)
{
    int rc = check(e);
    if (rc) return rc;
    if (E) CK(cudaMemcpyAsync(E, e->E.p, e->C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (naccept) CK(cudaMemcpyAsync(naccept, e->nacc.p, e->C * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    if (ntrials) CK(cudaMemcpyAsync(ntrials, e->ntri.p, e->C * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

int smcb_set_chain_energy(smcb_engine *e, const double *E)
{
    int rc = check(e);
    if (rc) return rc;
    if (!E) return fail(SMCB_ERR_ARG, "E is null");
    CK(cudaMemcpyAsync(e->E.p, E, e->C * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->energy_valid = true;
    return SMCB_OK;
}

int smcb_reset_counters(smcb_engine *e)
{
    int rc = check(e);
    if (rc) return rc;
    CK(cudaMemsetAsync(e->nacc.p, 0, e->C * sizeof(long long), e->stream));
    CK(cudaMemsetAsync(e->ntri.p, 0, e->C * sizeof(long long), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

// -------------------------------------------------------------------- sweep
static int finish_timed(smcb_engine *e, int launches)
{
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaMemcpyAsync(e->last_pairs, e->pairs.p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    e->last_launches = launches;
    return SMCB_OK;
}

static int sweep_common(smcb_engine *e, int nsweeps, int mode, bool fed, const double *displ,
                        const int64_t *offset, const double *u, uint8_t *accepted,
                        double *E_trace = nullptr, int32_t *acc_trace = nullptr)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (nsweeps < 0) return fail(SMCB_ERR_ARG, "nsweeps < 0");
    if (mode == SMCB_FP32) return fail(SMCB_ERR_ARG, "SMCB_FP32 covers the static evaluation and the all-particle step; the sweep runs in SMCB_FAST or SMCB_STRICT");
    if (mode != SMCB_FAST && mode != SMCB_STRICT) return fail(SMCB_ERR_ARG, "bad mode %d", mode);
    if (e->N > kSweepBlockMaxN)
        return fail(SMCB_ERR_ARG, "sweep kernels support N <= %d (N = %d): use smcb_step_allparticle", kSweepBlockMaxN, e->N);
    if (nsweeps == 0) return SMCB_OK;
    // stale running energy: the FAST warp-per-chain kernels take it from their own cache rebuild, the others need an evaluation
    const bool kernel_refreshes = mode == SMCB_FAST && e->N <= kSweepMaxN;     // the warp-per-chain FAST kernels take E from their cache rebuild
    if (!e->energy_valid && !kernel_refreshes && (rc = refresh_energy(e, mode))) return rc;
    SweepArgs a{};
    a.refresh_E = (!e->energy_valid && kernel_refreshes) ? 1 : 0;
    a.nsweeps = nsweeps;
    a.rng = RngArgs{(uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->chain0, e->step};
    const size_t sc = (size_t)nsweeps * e->C;
    if (fed) {
        if (!displ || !offset || !u) return fail(SMCB_ERR_ARG, "fed sweep needs displ, offset and u");
        CK(e->fed_a.ensure(sc * 3 * e->N));
        CK(e->fed_b.ensure(sc * e->N));
        CK(e->fed_off.ensure(sc));
        CK(cudaMemcpyAsync(e->fed_a.p, displ, sc * 3 * e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->fed_b.p, u, sc * e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->fed_off.p, offset, sc * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
        a.displ = e->fed_a.p; a.u = e->fed_b.p; a.offset = e->fed_off.p;
        if (accepted) { CK(e->fed_acc.ensure(sc * e->N)); a.accepted = e->fed_acc.p; }
    }
    const bool traced = E_trace != nullptr || acc_trace != nullptr;
    if (traced) {
        CK(e->trace_E.ensure(sc)); CK(e->trace_acc.ensure(sc));
        a.trace_E = e->trace_E.p; a.trace_acc = e->trace_acc.p;
    }
    if (e->capture_cache && mode == SMCB_FAST) {
        CK(e->cache_out.ensure((size_t)e->C * 5 * e->Npad));
        a.cache_out = e->cache_out.p;
    }
    CK(cudaMemsetAsync(e->pairs.p, 0, 16 * sizeof(unsigned long long), e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    const DevChains d = e->chains();
    a.dense_hint = e->sweep_dense;
    CK(mode == SMCB_STRICT ? launch_sweep_strict(fed, d, a, e->stream) : launch_sweep_fast(fed, d, a, e->stream));
    if ((rc = finish_timed(e, 1))) return rc;
    if (mode == SMCB_FAST) e->sweep_dense = e->last_pairs[1] * 50ull > e->last_pairs[0] ? 1 : 0;
    if (fed && accepted) {
        CK(cudaMemcpyAsync(accepted, e->fed_acc.p, sc * e->N, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    if (traced) {
        if (E_trace) CK(cudaMemcpyAsync(E_trace, e->trace_E.p, sc * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (acc_trace) CK(cudaMemcpyAsync(acc_trace, e->trace_acc.p, sc * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    if (!fed) e->step += (uint64_t)nsweeps;
    e->forces_valid = false;
    e->energy_valid = true;
    return SMCB_OK;
}

int smcb_sweep_fed(smcb_engine *e, int nsweeps, int mode, const double *displ, const int64_t *offset,
                   const double *u, uint8_t *accepted)
{
    return sweep_common(e, nsweeps, mode, true, displ, offset, u, accepted);
}

int smcb_sweep(smcb_engine *e, int nsweeps, int mode)
{
    return sweep_common(e, nsweeps, mode, false, nullptr, nullptr, nullptr, nullptr);
}

int smcb_sweep_traced(smcb_engine *e, int nsweeps, int mode, const double *displ, const int64_t *offset,
                      const double *u, double *E_trace, int32_t *acc_trace)
{
    const bool fed = displ || offset || u;
    return sweep_common(e, nsweeps, mode, fed, displ, offset, u, nullptr, E_trace, acc_trace);
}

}  // extern "C"

// ------------------------------------------- the host-buffer step, pipelined
// One call = upload the chains' positions from the caller's HOST buffer, advance them, (optionally) gather the
// observables, download positions and chain state.  The batch is cut into kParts blocks of chains, each on its own
// stream: block p's PCIe copies run while the other blocks' kernels do, so the call costs max(kernels, copies) plus
// one block's copies instead of their sum.  Page-locked (pinned) host buffers are needed for the overlap; pageable
// ones work, serialised by the driver.  Results are those of smcb_set_positions + smcb_sweep (or
// smcb_step_allparticle) + smcb_gather + smcb_get_positions + smcb_get_chain_state on the whole batch: chains are
// independent and a chain's random stream is a function of its global id only.
static DevChains sub_chains(smcb_engine *e, int c0, int cn)
{
    DevChains d = e->chains();
    d.C = cn;
    d.pos += (size_t)c0 * 3 * e->Npad;
    d.E += c0; d.nacc += c0; d.ntri += c0;
    d.extent += (size_t)2 * c0;
    if (d.nparams != 1) d.params += c0;
    return d;
}

extern "C" int smcb_sweep_host(smcb_engine *e, double *R, int nsteps, int mode, int kernel, int gather,
                               double *E_out, int64_t *naccept_out, int64_t *ntrials_out)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->have_params) return fail(SMCB_ERR_STATE, "smcb_set_params has not been called");
    if (!R) return fail(SMCB_ERR_ARG, "R is null");
    if (nsteps <= 0) return fail(SMCB_ERR_ARG, "nsteps must be positive");
    if (mode != SMCB_FAST && mode != SMCB_STRICT) return fail(SMCB_ERR_ARG, "smcb_sweep_host runs in SMCB_FAST or SMCB_STRICT (got mode %d)", mode);
    if (kernel != 0 && kernel != 1) return fail(SMCB_ERR_ARG, "kernel: 0 = sweep (oneParticleMoves), 1 = all-particle step");
    if (kernel == 0 && e->N > kSweepBlockMaxN) return fail(SMCB_ERR_ARG, "sweep kernels support N <= %d (N = %d)", kSweepBlockMaxN, e->N);
    if (kernel == 1 && (size_t)(6 * e->Npad + 256) * sizeof(double) > 227 * 1024) return fail(SMCB_ERR_ARG, "N = %d does not fit one CTA's shared memory", e->N);
    if (gather && !e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (gather && e->obs_reduced) return fail(SMCB_ERR_STATE, "the observable block holds an all-reduced total: smcb_obs_reset first");
    const int C = e->C, N = e->N, Npad = e->Npad;
    const size_t cn3 = (size_t)C * 3 * Npad;
    CK(e->stage.ensure((size_t)C * 3 * N));
    if (kernel == 1) { CK(e->F.ensure(cn3)); CK(e->Fn.ensure(cn3)); CK(e->dl.ensure(cn3)); }
    if (gather) CK(e->chain_mom.ensure((size_t)C * 5));
    // blocks of chains: enough of them to hide the copies (>= ~12 MB each way per block, a quarter of a millisecond of
    // PCIe), never so many that a block's kernels underfill the GPU or the launches outnumber the work
    int parts = (int)std::min<size_t>(smcb_engine::kParts, std::max<size_t>(1, ((size_t)C * 3 * N * sizeof(double)) / (12u << 20)));
    while (parts > 1 && C / parts < 256) parts /= 2;
    int eparts = 1;                                               // blocks per chain of the FAST evaluation, decided per block size
    {
        DevChains dd = sub_chains(e, 0, (C + parts - 1) / parts);
        eparts = evaluate_fast_parts(dd);
    }
    CK(e->eval_partials.ensure((size_t)C * eparts * kTot));
    if (e->eval_tickets.n < (size_t)C) {
        CK(e->eval_tickets.ensure(C));
        CK(cudaMemsetAsync(e->eval_tickets.p, 0, C * sizeof(unsigned), e->stream));
    }
    // an error return below must not leave the block streams running behind the caller's back (the next call on the
    // engine's own stream would not be ordered after them): drain the device on every early exit
    struct Drain {
        bool armed = true;
        ~Drain() { if (armed) cudaDeviceSynchronize(); }
    } drain;
    CK(cudaMemsetAsync(e->pairs.p, 0, 16 * sizeof(unsigned long long), e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    CK(cudaEventRecord(e->pstart, e->stream));
    GatherArgs g{};
    if (gather) {
        g.totals = e->totals.p; g.rbin = e->rbin.p; g.counters = e->counters.p; g.moments = e->moments.p;
        g.chain_mom = e->chain_mom.p; g.ngroups = e->ngroups;
        g.wall_virial_intended = e->wall_virial_intended ? 1 : 0;
        g.u64_per_group = e->u64_per_group(); g.f64_per_group = e->f64_per_group();
        g.nebins = e->nebins; g.e_lo = e->e_lo; g.e_hi = e->e_hi;
    }
    const bool kernel_refreshes = kernel == 1 || (mode == SMCB_FAST && N <= kSweepMaxN);     // no separate energy evaluation needed
    for (int p = 0; p < parts; p++) {
        const int c0 = (int)((long long)C * p / parts), c1 = (int)((long long)C * (p + 1) / parts), cn = c1 - c0;
        if (cn <= 0) continue;
        cudaStream_t st = e->pstream[p];
        CK(cudaStreamWaitEvent(st, e->pstart, 0));
        const size_t aoff = (size_t)c0 * 3 * N, an = (size_t)cn * 3 * N;
        DevChains d = sub_chains(e, c0, cn);
        // in: positions; E <- energy + wallsEnergy of them (SMC.c:48) comes out of the kernel's own cache rebuild
        CK(cudaMemcpyAsync(e->stage.p + aoff, R + aoff, an * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(launch_aos_to_soa(e->stage.p + aoff, d.pos, cn, N, Npad, 3, st));
        CK(launch_chain_extent(d, e->extent.p + (size_t)2 * c0, e->extent_flag.p, st));
        if (!kernel_refreshes) {             // STRICT / block-per-chain sweeps read d.E: evaluate it first
            EvalOut o{};
            o.totals = e->totals.p + (size_t)c0 * kTot;
            DevChains dev = d;
            dev.step_scale = 1.0; dev.pair_counts = nullptr;
            if (mode == SMCB_STRICT) CK(launch_evaluate_strict(dev, o, st));
            else CK(launch_evaluate_fast_screened(dev, o, eparts, e->eval_partials.p + (size_t)c0 * eparts * kTot, e->eval_tickets.p + c0, st));
            CK(launch_energy_from_totals(o.totals, d.E, cn, st));
        }
        const RngArgs rng{(uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->chain0 + (uint32_t)c0, e->step};
        if (kernel == 0) {
            SweepArgs a{};
            a.nsweeps = nsteps; a.rng = rng; a.dense_hint = e->sweep_dense; a.refresh_E = kernel_refreshes ? 1 : 0;
            CK(mode == SMCB_STRICT ? launch_sweep_strict(false, d, a, st) : launch_sweep_fast(false, d, a, st));
        } else {
            StepArgs a{};
            a.nsteps = nsteps; a.refresh = 1; a.rng = rng;
            a.F = e->F.p + (size_t)c0 * 3 * Npad; a.Fn = e->Fn.p + (size_t)c0 * 3 * Npad; a.dl = e->dl.p + (size_t)c0 * 3 * Npad;
            CK(mode == SMCB_STRICT ? launch_allparticle_strict(false, d, a, st) : launch_allparticle_fast(false, d, a, st));
        }
        CK(cudaEventRecord(e->pev[p], st));                  // the block's chains are advanced: the gather may read them
        CK(cudaStreamWaitEvent(e->stream, e->pev[p], 0));
        // out: positions and chain state
        CK(launch_soa_to_aos(d.pos, e->stage.p + aoff, cn, N, Npad, 3, st));
        CK(cudaMemcpyAsync(R + aoff, e->stage.p + aoff, an * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (E_out) CK(cudaMemcpyAsync(E_out + c0, d.E, cn * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (naccept_out) CK(cudaMemcpyAsync(naccept_out + c0, d.nacc, cn * sizeof(long long), cudaMemcpyDeviceToHost, st));
        if (ntrials_out) CK(cudaMemcpyAsync(ntrials_out + c0, d.ntri, cn * sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    if (gather) {
        // one evaluation + gather of the WHOLE batch once every block is advanced (a 512-thread evaluation block needs an SM
        // free of sweep blocks: interleaving it per block with the other blocks' sweeps drained SMs and cost 5 ms); the
        // blocks' downloads run meanwhile on their own streams
        EvalOut o{};
        o.totals = e->totals.p;
        if ((rc = run_evaluate(e, SMCB_FAST, o))) return rc;
        CK(launch_gather(e->chains(), g, e->stream));
    }
    for (int p = 0; p < parts; p++) {
        CK(cudaEventRecord(e->pev[p], e->pstream[p]));       // the block's downloads are done
        CK(cudaStreamWaitEvent(e->stream, e->pev[p], 0));
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaMemcpyAsync(e->pairs_pinned, e->pairs.p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->flag_pinned, e->extent_flag.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    drain.armed = false;                     // every block stream has been joined into the engine's stream
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    if (*e->flag_pinned) {
        CK(cudaMemsetAsync(e->extent_flag.p, 0, sizeof(int), e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->have_pos = false;
        return fail(SMCB_ERR_ARG, "positions contain NaN or coordinates more than 2^20 box lengths from the origin");
    }
    for (int k = 0; k < 16; k++) e->last_pairs[k] = e->pairs_pinned[k];
    e->last_launches = parts * 3 + (gather ? 3 : 0);
    if (kernel == 0 && mode == SMCB_FAST) e->sweep_dense = e->last_pairs[1] * 50ull > e->last_pairs[0] ? 1 : 0;
    e->step += (uint64_t)nsteps;
    e->have_pos = true; e->energy_valid = true; e->forces_valid = kernel == 1;
    if (gather) e->totals_valid = true;
    return SMCB_OK;
}

extern "C" {
// ------------------------------------------------------- all-particle step
static int step_common(smcb_engine *e, int nsteps, int mode, bool fed, const double *xi, const double *u,
                       double *lnap, uint8_t *accepted)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (nsteps < 0) return fail(SMCB_ERR_ARG, "nsteps < 0");
    if (mode != SMCB_FAST && mode != SMCB_STRICT && mode != SMCB_FP32) return fail(SMCB_ERR_ARG, "bad mode %d", mode);
    if ((size_t)(6 * e->Npad + 256) * sizeof(double) > 227 * 1024)
        return fail(SMCB_ERR_ARG, "N = %d does not fit one CTA's shared memory", e->N);
    if (nsteps == 0) return SMCB_OK;
    const size_t cn3 = (size_t)e->C * 3 * e->Npad;
    CK(e->F.ensure(cn3)); CK(e->Fn.ensure(cn3)); CK(e->dl.ensure(cn3));
    StepArgs a{};
    a.nsteps = nsteps;
    a.refresh = (e->forces_valid && e->energy_valid) ? 0 : 1;
    a.rng = RngArgs{(uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->chain0, e->step};
    a.F = e->F.p; a.Fn = e->Fn.p; a.dl = e->dl.p;
    const size_t sc = (size_t)nsteps * e->C;
    if (fed) {
        if (!xi || !u) return fail(SMCB_ERR_ARG, "fed step needs xi and u");
        CK(e->fed_a.ensure(sc * 3 * e->N));
        CK(e->fed_b.ensure(sc));
        CK(cudaMemcpyAsync(e->fed_a.p, xi, sc * 3 * e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->fed_b.p, u, sc * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        a.xi = e->fed_a.p; a.u = e->fed_b.p;
    }
    if (lnap) { CK(e->stage.ensure(sc > (size_t)e->C * 3 * e->N ? sc : (size_t)e->C * 3 * e->N)); a.lnap = e->stage.p; }
    if (accepted) { CK(e->fed_acc.ensure(sc)); a.accepted = e->fed_acc.p; }
    CK(cudaMemsetAsync(e->pairs.p, 0, 16 * sizeof(unsigned long long), e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    const DevChains d = e->chains();
    CK(mode == SMCB_STRICT ? launch_allparticle_strict(fed, d, a, e->stream)
                           : (mode == SMCB_FP32 ? launch_allparticle_f32(fed, d, a, e->stream) : launch_allparticle_fast(fed, d, a, e->stream)));
    if ((rc = finish_timed(e, 1))) return rc;
    if (lnap) CK(cudaMemcpyAsync(lnap, a.lnap, sc * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (accepted) CK(cudaMemcpyAsync(accepted, a.accepted, sc, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (!fed) e->step += (uint64_t)nsteps;
    e->forces_valid = mode != SMCB_FP32;        // the FP32 step leaves float images in the force buffers
    e->energy_valid = true;
    return SMCB_OK;
}

int smcb_step_allparticle_fed(smcb_engine *e, int nsteps, int mode, const double *xi, const double *u,
                              double *lnap, uint8_t *accepted)
{
    return step_common(e, nsteps, mode, true, xi, u, lnap, accepted);
}

int smcb_step_allparticle(smcb_engine *e, int nsteps, int mode)
{
    return step_common(e, nsteps, mode, false, nullptr, nullptr, nullptr, nullptr);
}

}  // extern "C"

// ------------------------------------------------------ step-size control
extern "C" int smcb_tune_step_size(smcb_engine *e, int kernel, int mode, double target, int rounds, int nsteps_per_round)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (kernel != 0 && kernel != 1) return fail(SMCB_ERR_ARG, "kernel: 0 = sweep, 1 = all-particle step");
    if (!(target > 0.0 && target < 1.0) || rounds <= 0 || nsteps_per_round <= 0)
        return fail(SMCB_ERR_ARG, "need 0 < target < 1, rounds > 0, nsteps_per_round > 0");
    if (e->nparams == 1 && e->C > 1) {              // one shared parameter set: give every chain its own copy
        smcb_chain_params one;
        CK(cudaMemcpyAsync(&one, e->params.p, sizeof one, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        std::vector<smcb_chain_params> all((size_t)e->C, one);
        CK(e->params.ensure(e->C));
        CK(cudaMemcpyAsync(e->params.p, all.data(), all.size() * sizeof one, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->nparams = e->C;
    }
    for (int r = 0; r < rounds; r++) {
        CK(cudaMemsetAsync(e->nacc.p, 0, e->C * sizeof(long long), e->stream));
        CK(cudaMemsetAsync(e->ntri.p, 0, e->C * sizeof(long long), e->stream));
        rc = kernel == 0 ? smcb_sweep(e, nsteps_per_round, mode) : smcb_step_allparticle(e, nsteps_per_round, mode);
        if (rc) return rc;
        const double gain = std::max(1.0, 3.0 / (1.0 + 0.25 * r));   // decreasing, then constant: A settles where acceptance = target
        CK(launch_adapt_step(e->params.p, e->nacc.p, e->ntri.p, e->C, target, gain, 1e-14, 1e3, e->stream));
        e->forces_valid = e->forces_valid && kernel == 1;
    }
    std::vector<smcb_chain_params> all((size_t)e->nparams);
    CK(cudaMemcpyAsync(all.data(), e->params.p, all.size() * sizeof(smcb_chain_params), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    uint64_t h = fnv1a(all.data(), all.size() * sizeof(smcb_chain_params));
    if (e->nwalls > 0) {
        std::vector<double> W((size_t)e->nwalls * 2 * e->M * e->M);
        CK(cudaMemcpyAsync(W.data(), e->W.p, W.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        h = fnv1a(W.data(), W.size() * sizeof(double), h);
    }
    e->params_hash = h;
    return SMCB_OK;
}

extern "C" int smcb_get_step_sizes(smcb_engine *e, double *A)
{
    int rc = check(e);
    if (rc) return rc;
    if (!A) return fail(SMCB_ERR_ARG, "A is null");
    if (!e->have_params) return fail(SMCB_ERR_STATE, "smcb_set_params has not been called");
    std::vector<smcb_chain_params> all((size_t)e->nparams);
    CK(cudaMemcpyAsync(all.data(), e->params.p, all.size() * sizeof(smcb_chain_params), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < e->C; c++) A[c] = all[e->nparams == 1 ? 0 : c].A;
    return SMCB_OK;
}

extern "C" {
// -------------------------------------------------------------- observables
int smcb_obs_configure(smcb_engine *e, int nebins, double e_lo, double e_hi)
{
    int rc = check(e);
    if (rc) return rc;
    if (nebins <= 0 || !(e_hi > e_lo)) return fail(SMCB_ERR_ARG, "need nebins>0 and e_hi>e_lo");
    e->nebins = nebins; e->e_lo = e_lo; e->e_hi = e_hi;
    return obs_alloc(e, true);
}

int smcb_obs_layout_get(smcb_engine *e, smcb_obs_layout *out)
{
    int rc = check(e);
    if (rc) return rc;
    if (!out) return fail(SMCB_ERR_ARG, "out is null");
    out->ngroups = e->ngroups;
    out->nvox = SMCB_NCX * SMCB_NCX * SMCB_NCZ;
    out->nz = SMCB_NCZ;
    out->nebins = e->nebins;
    out->e_lo = e->e_lo; out->e_hi = e->e_hi;
    out->u64_per_group = e->u64_per_group();
    out->f64_per_group = e->f64_per_group();
    out->u64_total = e->u64_per_group() * e->ngroups;
    out->f64_total = e->f64_per_group() * e->ngroups;
    return SMCB_OK;
}

int smcb_gather(smcb_engine *e)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (!e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (e->obs_reduced)
        return fail(SMCB_ERR_STATE, "the observable block holds an all-reduced total (smcb_obs_allreduce / smcb_obs_import_device): "
                                    "gathering on top of it would count the other ranks' samples again at the next reduce; call smcb_obs_reset first");
    EvalOut o{};
    o.totals = e->totals.p;
    CK(cudaEventRecord(e->ev0, e->stream));
    if ((rc = run_evaluate(e, SMCB_FAST, o))) return rc;
    GatherArgs g{};
    CK(e->chain_mom.ensure((size_t)e->C * 5));
    g.totals = e->totals.p; g.rbin = e->rbin.p; g.counters = e->counters.p; g.moments = e->moments.p;
    g.chain_mom = e->chain_mom.p; g.ngroups = e->ngroups;
    g.wall_virial_intended = e->wall_virial_intended ? 1 : 0;
    g.u64_per_group = e->u64_per_group(); g.f64_per_group = e->f64_per_group();
    g.nebins = e->nebins; g.e_lo = e->e_lo; g.e_hi = e->e_hi;
    CK(launch_gather(e->chains(), g, e->stream));
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    e->last_launches = 3;
    e->totals_valid = true;
    return SMCB_OK;
}

int smcb_obs_reset(smcb_engine *e)
{
    int rc = check(e);
    if (rc) return rc;
    return obs_alloc(e, false);
}

int smcb_obs_get(smcb_engine *e, uint64_t *counters, double *moments)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (counters) CK(cudaMemcpyAsync(counters, e->counters.p, e->u64_per_group() * e->ngroups * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    if (moments) CK(cudaMemcpyAsync(moments, e->moments.p, e->f64_per_group() * e->ngroups * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

int smcb_obs_export_device(smcb_engine *e, void *counters_dev, void *moments_dev)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (counters_dev) CK(cudaMemcpyAsync(counters_dev, e->counters.p, e->u64_per_group() * e->ngroups * sizeof(uint64_t), cudaMemcpyDeviceToDevice, e->stream));
    if (moments_dev) CK(cudaMemcpyAsync(moments_dev, e->moments.p, e->f64_per_group() * e->ngroups * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

// The delta-reduce step without a host round trip: the block is copied out and the accumulators are zeroed (Rbin kept)
// on the engine's stream; the call returns at once.  The caller orders its collective after the engine's stream (an event
// recorded on smcb_stream(e)) and may launch the next sweep immediately: the all-reduce then runs under it.
int smcb_obs_export_reset_async(smcb_engine *e, void *counters_dev, void *moments_dev)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (!counters_dev || !moments_dev) return fail(SMCB_ERR_ARG, "null device buffer");
    const size_t nc = e->u64_per_group() * e->ngroups * sizeof(uint64_t), nm = e->f64_per_group() * e->ngroups * sizeof(double);
    CK(cudaMemcpyAsync(counters_dev, e->counters.p, nc, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(moments_dev, e->moments.p, nm, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemsetAsync(e->counters.p, 0, nc, e->stream));
    CK(cudaMemsetAsync(e->moments.p, 0, nm, e->stream));
    e->obs_reduced = false;
    return SMCB_OK;
}

int smcb_obs_import_device(smcb_engine *e, const void *counters_dev, const void *moments_dev)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->counters.p) return fail(SMCB_ERR_STATE, "observable block not allocated");
    if (counters_dev) CK(cudaMemcpyAsync(e->counters.p, counters_dev, e->u64_per_group() * e->ngroups * sizeof(uint64_t), cudaMemcpyDeviceToDevice, e->stream));
    if (moments_dev) CK(cudaMemcpyAsync(e->moments.p, moments_dev, e->f64_per_group() * e->ngroups * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

}  // extern "C"

// ---- single-process multi-GPU all-reduce of the observable blocks (NCCL loaded at run time) ----------
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString;
    }
};
NcclApi g_nccl;
std::vector<int> g_nccl_devs;          // device list the cached communicators were built for
std::vector<void *> g_nccl_comms;
constexpr int kNcclSum = 0, kNcclUint64 = 5, kNcclFloat64 = 8;     // nccl.h: ncclSum, ncclUint64, ncclFloat64
}  // namespace

static std::mutex g_nccl_mutex;            // guards the communicator cache (g_nccl_comms / g_nccl_devs)

extern "C" int smcb_obs_allreduce(smcb_engine **engines, int n)
{
    if (!engines || n <= 0) return fail(SMCB_ERR_ARG, "need n > 0 engines");
    for (int i = 0; i < n; i++) {
        smcb_engine *e = engines[i];
        if (!e || !e->counters.p) return fail(SMCB_ERR_STATE, "engine %d has no observable block", i);
        if (e->ngroups != engines[0]->ngroups || e->nebins != engines[0]->nebins || e->u64_per_group() != engines[0]->u64_per_group() ||
            e->f64_per_group() != engines[0]->f64_per_group() || e->e_lo != engines[0]->e_lo || e->e_hi != engines[0]->e_hi)
            return fail(SMCB_ERR_ARG, "engine %d has a different observable layout (groups, energy bins or range)", i);
        if (e->obs_reduced) return fail(SMCB_ERR_STATE, "engine %d already holds an all-reduced block: smcb_obs_reset before reducing again", i);
        for (int j = 0; j < i; j++)
            if (engines[j]->device == e->device) return fail(SMCB_ERR_ARG, "engines %d and %d share GPU %d", j, i, e->device);
    }
    if (n == 1) return SMCB_OK;
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!g_nccl.load()) return fail(SMCB_ERR_STATE, "NCCL not available (dlopen libnccl.so.2: %s)", dlerror());
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) devs[i] = engines[i]->device;
    if (devs != g_nccl_devs) {
        for (void *c : g_nccl_comms) g_nccl.CommDestroy(c);
        g_nccl_comms.assign(n, nullptr);
        const int rc = g_nccl.CommInitAll(g_nccl_comms.data(), n, devs.data());
        if (rc != 0) { g_nccl_comms.clear(); g_nccl_devs.clear(); cudaSetDevice(prev_dev); return fail(SMCB_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(rc)); }
        g_nccl_devs = devs;
    }
    const size_t ncnt = engines[0]->u64_per_group() * engines[0]->ngroups, nmom = engines[0]->f64_per_group() * engines[0]->ngroups;
    int rc = g_nccl.GroupStart();
    cudaError_t cerr = cudaSuccess;
    for (int i = 0; i < n && rc == 0 && cerr == cudaSuccess; i++) {          // an error leaves the loop, never the open group
        smcb_engine *e = engines[i];
        cerr = cudaSetDevice(e->device);
        if (cerr != cudaSuccess) break;
        rc = g_nccl.AllReduce(e->counters.p, e->counters.p, ncnt, kNcclUint64, kNcclSum, g_nccl_comms[i], e->stream);
        if (rc == 0) rc = g_nccl.AllReduce(e->moments.p, e->moments.p, nmom, kNcclFloat64, kNcclSum, g_nccl_comms[i], e->stream);
    }
    const int rc2 = g_nccl.GroupEnd();
    for (int i = 0; i < n && cerr == cudaSuccess; i++) {
        cerr = cudaSetDevice(engines[i]->device);
        if (cerr == cudaSuccess) cerr = cudaStreamSynchronize(engines[i]->stream);
    }
    cudaSetDevice(prev_dev);
    if (rc != 0 || rc2 != 0) return fail(SMCB_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(rc != 0 ? rc : rc2));
    if (cerr != cudaSuccess) return fail(SMCB_ERR_CUDA, "smcb_obs_allreduce: %s", cudaGetErrorString(cerr));
    for (int i = 0; i < n; i++) engines[i]->obs_reduced = true;     // every block now holds the job's totals
    return SMCB_OK;
}

// release the cached NCCL communicators (optional; call when no engine will all-reduce again)
extern "C" int smcb_obs_allreduce_teardown(void)
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    for (void *c : g_nccl_comms) if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    g_nccl_comms.clear();
    g_nccl_devs.clear();
    return SMCB_OK;
}

extern "C" {
int smcb_get_rbin(smcb_engine *e, int32_t *rbin)
{
    int rc = check(e);
    if (rc) return rc;
    if (!rbin) return fail(SMCB_ERR_ARG, "rbin is null");
    CK(cudaMemcpyAsync(rbin, e->rbin.p, (size_t)e->C * e->N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

int smcb_set_rbin(smcb_engine *e, const int32_t *rbin)
{
    int rc = check(e);
    if (rc) return rc;
    if (!rbin) return fail(SMCB_ERR_ARG, "rbin is null");
    CK(cudaMemcpyAsync(e->rbin.p, rbin, (size_t)e->C * e->N * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SMCB_OK;
}

// ---- the wall virial as the reference meant it -------------------------------
int smcb_get_wall_virial(smcb_engine *e, double *vir_wall)
{
    int rc = check(e);
    if (rc) return rc;
    if (!vir_wall) return fail(SMCB_ERR_ARG, "vir_wall is null");
    if (!e->totals_valid) return fail(SMCB_ERR_STATE, "call smcb_evaluate or smcb_gather first");
    std::vector<double> t((size_t)kTot * e->C);
    CK(cudaMemcpyAsync(t.data(), e->totals.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < e->C; c++) vir_wall[c] = t[kTot * c + 4];
    return SMCB_OK;
}

int smcb_obs_set_wall_virial(smcb_engine *e, int intended)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    e->wall_virial_intended = intended != 0;
    return SMCB_OK;
}

// --------------------------------------------------------- checkpoint / resume

int smcb_checkpoint_save(smcb_engine *e, const char *path)
{
    int rc = need_ready(e);
    if (rc) return rc;
    if (!path) return fail(SMCB_ERR_ARG, "path is null");
    if (!e->energy_valid && (rc = refresh_energy(e, SMCB_FAST))) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return fail(SMCB_ERR_ARG, "cannot open %s for writing", path);
    CkptHeader h{};
    memcpy(h.magic, "SMCB200", 8);
    h.version = 2; h.params_hash = e->params_hash; h.C = e->C; h.N = e->N; h.M = e->M; h.ngroups = e->ngroups; h.nebins = e->nebins;
    h.chain0 = e->chain0; h.pad_ = (uint32_t)e->sweep_dense; h.seed = e->seed; h.step = e->step; h.step_scale = e->step_scale;
    h.e_lo = e->e_lo; h.e_hi = e->e_hi;
    h.n_counters = e->u64_per_group() * e->ngroups; h.n_moments = e->f64_per_group() * e->ngroups;
    std::vector<unsigned char> buf;
    int w = fwrite(&h, sizeof h, 1, f) == 1 ? 0 : -2;
    if (!w) w = put(f, e->pos.p, (size_t)e->C * 3 * e->Npad, e->stream, buf);
    if (!w) w = put(f, e->E.p, (size_t)e->C, e->stream, buf);
    if (!w) w = put(f, e->nacc.p, (size_t)e->C, e->stream, buf);
    if (!w) w = put(f, e->ntri.p, (size_t)e->C, e->stream, buf);
    if (!w) w = put(f, e->rbin.p, (size_t)e->C * e->N, e->stream, buf);
    if (!w) w = put(f, e->counters.p, (size_t)h.n_counters, e->stream, buf);
    if (!w) w = put(f, e->moments.p, (size_t)h.n_moments, e->stream, buf);
    if (fclose(f) != 0 && !w) w = -2;
    if (w == -1) return fail(SMCB_ERR_CUDA, "checkpoint_save: device copy failed");
    if (w) return fail(SMCB_ERR_ARG, "checkpoint_save: short write to %s", path);
    return SMCB_OK;
}

int smcb_checkpoint_load(smcb_engine *e, const char *path)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->have_params) return fail(SMCB_ERR_STATE, "smcb_set_params must be called before smcb_checkpoint_load");
    if (!path) return fail(SMCB_ERR_ARG, "path is null");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(SMCB_ERR_ARG, "cannot open %s", path);
    CkptHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SMCB200", 8) != 0 || h.version != 2) {
        fclose(f);
        return fail(SMCB_ERR_ARG, "%s is not a smcb200 checkpoint (version 2)", path);
    }
    if ((int)h.C != e->C || (int)h.N != e->N || (int)h.M != e->M || (int)h.ngroups != e->ngroups) {
        fclose(f);
        return fail(SMCB_ERR_ARG, "checkpoint is for %u chains x N=%u, M=%u, %u groups; engine has %d x N=%d, M=%d, %d groups",
                    h.C, h.N, h.M, h.ngroups, e->C, e->N, e->M, e->ngroups);
    }
    if (h.params_hash != e->params_hash) {
        fclose(f);
        return fail(SMCB_ERR_ARG, "checkpoint was written under different chain parameters / wall tables than smcb_set_params gave this engine");
    }
    // every size is recomputed from the engine's own shape; nothing from the file sizes a copy
    if (h.nebins == 0 || h.nebins > (1u << 20) || !(h.e_hi > h.e_lo)) {
        fclose(f);
        return fail(SMCB_ERR_ARG, "checkpoint header: bad energy histogram (%u bins)", h.nebins);
    }
    const size_t u64pg = (size_t)2 * SMCB_NCX * SMCB_NCX * SMCB_NCZ + SMCB_NCZ + h.nebins + 1;
    const size_t n_counters = u64pg * e->ngroups, n_moments = e->f64_per_group() * e->ngroups;
    if (h.n_counters != n_counters || h.n_moments != n_moments) {
        fclose(f);
        return fail(SMCB_ERR_ARG, "checkpoint header: observable block of %llu + %llu elements, expected %zu + %zu",
                    (unsigned long long)h.n_counters, (unsigned long long)h.n_moments, n_counters, n_moments);
    }
    // read the whole body into host memory first: a truncated file changes nothing on the device
    const size_t C = e->C, N = e->N, Npad = e->Npad;
    const size_t sz[7] = {C * 3 * Npad * sizeof(double), C * sizeof(double), C * sizeof(long long), C * sizeof(long long),
                          C * N * sizeof(int), n_counters * sizeof(unsigned long long), n_moments * sizeof(double)};
    size_t total = 0;
    for (size_t v : sz) total += v;
    std::vector<unsigned char> body(total);
    const size_t got = fread(body.data(), 1, total, f);
    const bool extra = got == total && fgetc(f) != EOF;
    fclose(f);
    if (got != total) return fail(SMCB_ERR_ARG, "checkpoint_load: %s is truncated (%zu of %zu body bytes)", path, got, total);
    if (extra) return fail(SMCB_ERR_ARG, "checkpoint_load: %s has trailing bytes", path);
    if ((int)h.nebins != e->nebins || h.e_lo != e->e_lo || h.e_hi != e->e_hi) {
        e->nebins = (int)h.nebins; e->e_lo = h.e_lo; e->e_hi = h.e_hi;
        if ((rc = obs_alloc(e, true))) return rc;
    }
    void *dst[7] = {e->pos.p, e->E.p, e->nacc.p, e->ntri.p, e->rbin.p, e->counters.p, e->moments.p};
    size_t off = 0;
    for (int k = 0; k < 7; k++) {
        CK(cudaMemcpyAsync(dst[k], body.data() + off, sz[k], cudaMemcpyHostToDevice, e->stream));
        off += sz[k];
    }
    CK(cudaStreamSynchronize(e->stream));
    e->obs_reduced = false;
    e->seed = h.seed; e->chain0 = h.chain0; e->step = h.step; e->step_scale = h.step_scale;
    e->sweep_dense = h.pad_ == 1u ? 1 : 0;      // the resumed run launches the kernel the saved run would have launched next
    e->have_pos = true; e->energy_valid = true; e->forces_valid = false;
    return positions_uploaded(e);
}

// -------------------------------------------------------------- measurement
int smcb_last_kernel_ms(smcb_engine *e, float *ms, int *launches)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    if (ms) *ms = e->last_ms;
    if (launches) *launches = e->last_launches;
    return SMCB_OK;
}

int smcb_last_pair_counts(smcb_engine *e, uint64_t *pairs_total, uint64_t *pairs_in_cutoff)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    if (pairs_total) *pairs_total = e->last_pairs[0];
    if (pairs_in_cutoff) *pairs_in_cutoff = e->last_pairs[1];
    return SMCB_OK;
}

int smcb_last_pair_tests(smcb_engine *e, uint64_t *pair_tests_executed)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    if (pair_tests_executed) *pair_tests_executed = e->last_pairs[2];
    return SMCB_OK;
}

// path statistics of k_sweep_spec (12 counters; all zero unless the library was built with -DSMCB_SPEC_STATS)
int smcb_debug_sweep_stats(smcb_engine *e, uint64_t *stats12)
{
    if (!e || !stats12) return fail(SMCB_ERR_ARG, "null argument");
    for (int k = 0; k < 12; k++) stats12[k] = e->last_pairs[3 + k];
    return SMCB_OK;
}

int smcb_measure_fp64_peak(smcb_engine *e, double *tflops, float *ms_out)
{
    int rc = check(e);
    if (rc) return rc;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, e->device));
    CK(e->peak_out.ensure(1));
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 2048;
    CK(launch_dfma_peak(e->peak_out.p, blocks, threads, 64, e->stream));       // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e->ev0, e->stream));
        CK(launch_dfma_peak(e->peak_out.p, blocks, threads, iters, e->stream));
        CK(cudaEventRecord(e->ev1, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        float ms;
        CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        if (ms < best) best = ms;
    }
    const double flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return SMCB_OK;
}

int smcb_device_positions(smcb_engine *e, void **ptr, size_t *bytes, int *npad)
{
    int rc = check(e);
    if (rc) return rc;
    if (ptr) *ptr = e->pos.p;
    if (bytes) *bytes = (size_t)e->C * 3 * e->Npad * sizeof(double);
    if (npad) *npad = e->Npad;
    return SMCB_OK;
}

int smcb_debug_capture_cache(smcb_engine *e, int on)
{
    if (!e) return fail(SMCB_ERR_ARG, "null engine");
    e->capture_cache = on != 0;
    return SMCB_OK;
}

int smcb_debug_get_cache(smcb_engine *e, double *e_tot, double *f_tot, double *nb)
{
    int rc = check(e);
    if (rc) return rc;
    if (!e->cache_out.p) return fail(SMCB_ERR_STATE, "no cache captured (smcb_debug_capture_cache + a FAST sweep first)");
    const size_t CN = (size_t)e->C * e->N;
    std::vector<double> h((size_t)e->C * 5 * e->Npad);
    CK(cudaMemcpyAsync(h.data(), e->cache_out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (size_t c = 0; c < (size_t)e->C; c++) {
        const double *q = h.data() + c * 5 * e->Npad;
        for (int j = 0; j < e->N; j++) {
            if (e_tot) e_tot[c * e->N + j] = q[j];
            if (f_tot) for (int k = 0; k < 3; k++) f_tot[(c * e->N + j) * 3 + k] = q[(1 + k) * e->Npad + j];
            if (nb) nb[c * e->N + j] = q[4 * e->Npad + j];
        }
    }
    (void)CN;
    return SMCB_OK;
}

void *smcb_stream(smcb_engine *e) { return e ? (void *)e->stream : nullptr; }

}  // extern "C"
