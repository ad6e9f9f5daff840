/* examples/batched_driver.c — a plain C driver on the batched C ABI (include/smcb200.h), the shape an
 * MPI-style or multi-GPU replacement of the reference's main.c takes.
 *
 * It reuses the reference's own API from the drop-in for everything main.c does before the simulation
 * (initializeWalls, initializeBox, the N / M / a0 / b0 / LJ_CUTOFF macros of SMC.h), then advances
 * `chains` independent chains per GPU on every visible GPU (at most `maxgpus`), all-reduces the
 * observable blocks with smcb_obs_allreduce and prints the z density profile and the moments.
 *
 *   gcc -std=gnu11 -O2 -I montecarlo-surfacer_b200/dropin -I include examples/batched_driver.c \
 *       -L montecarlo-surfacer_b200 -lsmcb200 -lm -o batched_driver
 *   ./batched_driver [chains_per_gpu=256] [eqsteps=200] [maxsteps=400] [gather_lapse=40] [T=1.1] [maxgpus=8]
 */
#include "SMC.c"            /* the drop-in: SMC.h macros + initializeBox / initializeWalls */
#include "smcb200.h"

#define CHECK(call) do { if ((call) != SMCB_OK) { fprintf(stderr, "%s: %s\n", #call, smcb_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int chains = argc > 1 ? atoi(argv[1]) : 256, eqsteps = argc > 2 ? atoi(argv[2]) : 200;
    const int maxsteps = argc > 3 ? atoi(argv[3]) : 400, lapse = argc > 4 ? atoi(argv[4]) : 40;
    const double T = argc > 5 ? atof(argv[5]) : 1.1;
    const int maxgpus = argc > 6 ? atoi(argv[6]) : 8;
    const double L = 33, Lz = N < 150 ? 200 : 240;                 /* main.c:35-44 */

    double W[2 * M * M], R0[3 * N];
    initializeWalls(1.6, 0.0, 3.0, 0.5, W, NULL);                   /* main.c:74-87 (no CSV here) */
    initializeBox(L, Lz, N, R0);                                    /* main.c:112 */

    smcb_engine *eng[8];
    int ngpu = 0;
    for (; ngpu < maxgpus && ngpu < 8; ngpu++)
        if (smcb_create(&eng[ngpu], ngpu, chains, N, M) != SMCB_OK) break;     /* stops at the first missing device */
    if (ngpu == 0) { fprintf(stderr, "no GPU: %s\n", smcb_last_error()); return 1; }

    smcb_chain_params p = { .L = L, .Lz = Lz, .T = T, .A = 1.0 * T, .rc2 = LJ_CUTOFF * LJ_CUTOFF,
                            .zwall_a = a0, .zwall_b = b0, .flags = SMCB_WALL };
    for (int g = 0; g < ngpu; g++) {
        CHECK(smcb_set_params(eng[g], &p, 1, W, 1, 1));
        CHECK(smcb_broadcast_positions(eng[g], R0));
        CHECK(smcb_set_rng(eng[g], 20261018u, (uint32_t)(g * chains), 0));   /* global chain ids: disjoint streams */
        CHECK(smcb_set_step_scale(eng[g], 2.0));                              /* thermalisation with 2A, SMC.c:110 */
        CHECK(smcb_sweep(eng[g], eqsteps, SMCB_FAST));
        CHECK(smcb_set_step_scale(eng[g], 1.0));
        CHECK(smcb_reset_counters(eng[g]));
        CHECK(smcb_obs_reset(eng[g]));
    }
    int gathers = 0;
    for (int n = lapse; n <= maxsteps; n += lapse, gathers++)
        for (int g = 0; g < ngpu; g++) {                                      /* (a real driver gives each GPU a host thread) */
            CHECK(smcb_sweep(eng[g], lapse, SMCB_FAST));
            CHECK(smcb_gather(eng[g]));
        }
    CHECK(smcb_obs_allreduce(eng, ngpu));                                     /* the path's only collective (NCCL) */

    smcb_obs_layout lay;
    CHECK(smcb_obs_layout_get(eng[0], &lay));
    uint64_t *cnt = malloc(lay.u64_total * sizeof *cnt);
    double *mom = malloc(lay.f64_total * sizeof *mom);
    CHECK(smcb_obs_get(eng[0], cnt, mom));
    const uint64_t *zprof = cnt + 2 * (size_t)lay.nvox, nsamp = cnt[2 * (size_t)lay.nvox + lay.nz + lay.nebins];
    printf("gpus %d chains %d N %d gathers %d samples %llu\n", ngpu, ngpu * chains, N, gathers, (unsigned long long)nsamp);
    printf("mean_E %.6f mean_P %.6e acceptance %.4f\n", mom[0] / (double)nsamp + 3 * N * T / 2, mom[2] / (double)nsamp, mom[4] / (double)nsamp);
    printf("zprofile");
    uint64_t mass = 0;
    for (int k = 0; k < lay.nz; k++) { printf(" %llu", (unsigned long long)zprof[k]); mass += zprof[k]; }
    printf("\nmass %llu expected %llu\n", (unsigned long long)mass, (unsigned long long)nsamp * N);
    free(cnt); free(mom);

    /* A driver that keeps its configurations in HOST memory (as sMC keeps R, SMC.c:44, 117, 195) advances them with one
       call per step: upload, E <- energy + wallsEnergy, `lapse` sweeps, one gather, download - the chain blocks' copies
       overlap each other's kernels inside.  The all-reduced block above is final: reset it before gathering again. */
    double *R = malloc((size_t)chains * 3 * N * sizeof *R), *Ec = malloc((size_t)chains * sizeof *Ec);
    int64_t *na = malloc((size_t)chains * sizeof *na), *nt = malloc((size_t)chains * sizeof *nt);
    CHECK(smcb_get_positions(eng[0], R));
    CHECK(smcb_obs_reset(eng[0]));
    CHECK(smcb_reset_counters(eng[0]));
    CHECK(smcb_sweep_host(eng[0], R, lapse, SMCB_FAST, /*kernel: the sweep*/ 0, /*gather*/ 1, Ec, na, nt));
    long long acc = 0, tri = 0;
    for (int c = 0; c < chains; c++) { acc += na[c]; tri += nt[c]; }
    printf("host_step chains %d sweeps %d acceptance %.4f E0 %.6f\n", chains, lapse, tri ? (double)acc / (double)tri : 0.0, Ec[0]);
    const int host_ok = tri == (long long)chains * lapse * N;
    free(R); free(Ec); free(na); free(nt);
    for (int g = 0; g < ngpu; g++) smcb_destroy(eng[g]);
    return mass == nsamp * (uint64_t)N && host_ok ? 0 : 2;
}
