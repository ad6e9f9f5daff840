/* smcb200.h — C ABI of the B200-native Smart-Monte-Carlo engine (libsmcb200.so).
 *
 * This is the drop-in boundary for the hot path of Kryohi/MonteCarlo-Surfacer:
 * the force-biased single-particle sweep `oneParticleMoves` (reference
 * SMC.c:278-351) and the energy / force / wall / observable routines it and
 * `sMC` call (SMC.c:557-895, 912-927; prototypes SMC.h:92-117).  The reference
 * fixes N and M with macros and runs ONE chain per call; this ABI takes runtime
 * N, M and a batch of independent chains that advance in lock-step on one GPU.
 * One engine handle per GPU; chains shard across GPUs with no data-path
 * collective (only the observable block is all-reduced, see smcb_obs_*).
 *
 * Conventions (inherited from the reference where it has one):
 *   - plain C types, caller-owned HOST buffers unless a name says `_device`;
 *   - positions are AoS double[3N] per chain, chains concatenated (SMC.h:84);
 *   - W is (a,b) interleaved, m = i*M + j, 2*M*M doubles per wall table
 *     (SMC.c:475-501, 745-760);
 *   - every call returns 0 on success, <0 on error (smcb_last_error() has the
 *     text).  There is NO CPU fallback: without a CUDA device smcb_create fails.
 *   - calls on one handle are serialised by the caller; different handles are
 *     independent.  All calls are synchronous on return (smcb_sweep_host overlaps copies and kernels internally).
 *   - no identifier below collides with the reference's macros (N, M, a0, b0, rank, Ncx, Ncz,
 *     KMAX ...: SMC.h:26-61), so a translation unit built on SMC.h can include this header.
 */
#ifndef SMCB200_H
#define SMCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB_OK            0
#define SMCB_ERR_ARG      -1
#define SMCB_ERR_CUDA     -2
#define SMCB_ERR_NODEVICE -3
#define SMCB_ERR_STATE    -4

/* chain flags */
#define SMCB_WALL        1u  /* molecule-surface potential on (SMC.c:729-813)              */
#define SMCB_PERIODIC_Z  2u  /* bulk: z gets the minimum image with period Lz, proposals are
                                wrapped in z (SMC_noMPI_noWall.c:470-475); no reference in SMC.c */

/* arithmetic variants of the kernels */
#define SMCB_FAST    0  /* fused FMA formulation, one reciprocal per pair, tree reductions; the cutoff
                           test runs as a conservative packed-FP32 screen and every pair inside the
                           cutoff is evaluated in FP64 (same sums as an all-FP64 loop); <= 1e-12 rel.  */
#define SMCB_STRICT  1  /* the reference's IEEE operation order (no FMA, true divisions, sums in
                           the reference's particle order): bit-identical per-particle results   */
#define SMCB_FP32    2  /* optional single precision: smcb_evaluate, smcb_refresh_energy and smcb_step_allparticle(_fed)
                           only.  Coordinates are carried as float pairs (hi + lo), all pair arithmetic is FP32,
                           block sums FP64; energies and forces within 1e-5 of the reference's (relative to the
                           magnitude of what each sum adds up).  Molecules must be between the walls: the reference's
                           clamp for escaped ones (SMC.c:738-739) gives r^-12 = 1e48, beyond a float.              */

/* voxel grid of localDensityAndMobility (SMC.h:50-53: Ncx = Ncz = 33) */
#define SMCB_NCX 33
#define SMCB_NCZ 33

typedef struct smcb_engine smcb_engine;

/* Per-chain physical parameters.  The reference passes these as scalars
 * (L, Lz, T, A: SMC.h:92,102) or bakes them in as macros (LJ_CUTOFF, a0, b0:
 * SMC.h:32-39). */
typedef struct smcb_chain_params {
    double L;        /* x,y period                                                   */
    double Lz;       /* wall-to-wall distance (z period if SMCB_PERIODIC_Z)          */
    double T;        /* temperature                                                  */
    double A;        /* SMC step parameter: drift A/T * F, noise variance 2A         */
    double rc2;      /* LJ cutoff squared (9.0 in the reference)                     */
    double zwall_a, zwall_b;  /* flat z-wall 12-6 coefficients (the a0, b0 macros of SMC.h:32-33) */
    uint32_t flags;  /* SMCB_WALL | SMCB_PERIODIC_Z                                  */
    uint32_t wall;   /* index of this chain's wall table (0 .. nwalls-1)             */
    uint32_t group;  /* observable group (parameter-grid point) this chain adds to   */
    uint32_t pad_;
} smcb_chain_params;

/* ---- lifetime ----------------------------------------------------------- */
/* nchains chains of nparticles (the reference's N) molecules, wall tables of nsites_side^2 (M*M)
 * sites, on CUDA `device`. */
int smcb_create(smcb_engine **out, int device, int nchains, int nparticles, int nsites_side);
int smcb_destroy(smcb_engine *e);
const char *smcb_last_error(void);
/* library / device facts: "sm_100a", SM count, etc. (for logs and tests) */
int smcb_device_info(smcb_engine *e, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes);

/* ---- inputs ------------------------------------------------------------- */
/* nparams is nchains, or 1 to share one parameter set.  W holds nwalls tables
 * of 2*M*M doubles.  ngroups sizes the observable block. */
int smcb_set_params(smcb_engine *e, const smcb_chain_params *p, int nparams,
                    const double *W, int nwalls, int ngroups);
int smcb_set_positions(smcb_engine *e, const double *R);          /* nchains*3N, AoS */
int smcb_get_positions(smcb_engine *e, double *R);
/* replicate one configuration (3N doubles) into every chain */
int smcb_broadcast_positions(smcb_engine *e, const double *R0);
/* Philox4x32-10 stream identity: key = seed, counter carries (step, chain id,
 * particle).  chain0 is the global id of this engine's first chain so shards
 * of one job draw disjoint streams.  SMCB_STRICT kernels turn the bits into
 * Gaussians with a double-precision Box-Muller that the test oracle replays
 * exactly; SMCB_FAST kernels use a single-precision Box-Muller on the same
 * counter (csrc/philox.cuh) - a different, equally valid stream. */
int smcb_set_rng(smcb_engine *e, uint64_t seed, uint32_t chain0, uint64_t step0);

/* ---- static evaluation (rows a2-a10) ------------------------------------
 * For every chain: per-particle LJ energy e_lj[i] (= energySingle, SMC.c:557),
 * LJ force f_lj[3i..] (= forceSingle, SMC.c:589 = forces(), SMC.c:656),
 * per-particle surface energy e_wall[i] (= wallsEnergySingle, SMC.c:729) and
 * force f_wall[3i..] (= wallsForce, SMC.c:773), and the chain totals
 * U_lj (= energy, SMC.c:626), U_wall (= wallsEnergy, SMC.c:822), the LJ virial
 * sum vir_lj (pressure() = -vir_lj/(3 L^2 Lz), SMC.c:696) and the wall virial
 * sum AS THE REFERENCE WRITES IT vir_wall_ref (wallsPressure() =
 * -vir_wall_ref/(3 L^2 Lz), SMC.c:862-895, quirks kept).  Any output may be
 * NULL.  Per-particle arrays are nchains*N (energies) / nchains*3N (forces).
 * mode = SMCB_STRICT gives per-particle values bit-identical to the reference. */
int smcb_evaluate(smcb_engine *e, int mode,
                  double *e_lj, double *f_lj, double *e_wall, double *f_wall,
                  double *U_lj, double *U_wall, double *vir_lj, double *vir_wall_ref);

/* The wall virial as the reference MEANT it (its wallsPressure uses rz + L/2 instead of rz + Lz/2, has no
 * clamp and adds the flat-wall term once per in-cutoff site, SMC.c:880,888-889 - SURVEY.md App. B3): distance
 * to the nearer wall as in wallsEnergySingle, flat-wall term once, every site inside the cutoff.  Chain sums
 * computed by the last smcb_evaluate / smcb_gather; the corrected wall pressure is -vir_wall/(3 L^2 Lz). */
int smcb_get_wall_virial(smcb_engine *e, double *vir_wall);
/* which one enters the pressure moments of smcb_gather: 0 = the reference's arithmetic (default, what sMC
 * records, SMC.c:140), 1 = the corrected one */
int smcb_obs_set_wall_virial(smcb_engine *e, int intended);

/* ---- the sweep (row a1: oneParticleMoves, SMC.c:278-351) -----------------
 * nsweeps sweeps of N sequential single-particle Smart-MC trials per chain
 * (N <= 512: one warp per chain, SMCB_FAST or SMCB_STRICT; 512 < N <= 6016: one
 * block per chain, SMCB_FAST only; beyond that use the all-particle step).
 * Running energy (+= Un-Um on acceptance, SMC.c:341) and acceptance counts
 * accumulate in the chain state (smcb_get_chain_state).
 *
 * _fed: the random inputs the reference would have drawn are supplied by the
 * host (parity mode):  displ[s][c][3N] = vecBoxMuller(sqrt(2A),3N) (SMC.c:284),
 * offset[s][c] = the rand() of SMC.c:290, u[s][c][N] = rand()/RAND_MAX per
 * trial in visiting order (SMC.c:335).  accepted (nullable) receives one byte
 * per trial, [s][c][N] in visiting order.
 * Without _fed the engine draws from its Philox stream. */
int smcb_sweep_fed(smcb_engine *e, int nsweeps, int mode,
                   const double *displ, const int64_t *offset, const double *u,
                   uint8_t *accepted);
int smcb_sweep(smcb_engine *e, int nsweeps, int mode);
/* same, and returns what sMC records after every sweep (SMC.c:116-117, 194-195): the running
 * energy E_trace[s][c] (sMC's E[n+1]) and the accepted trials of that sweep acc_trace[s][c] (jj[n]).
 * displ/offset/u: all NULL (Philox) or all given (host-fed, as smcb_sweep_fed). */
int smcb_sweep_traced(smcb_engine *e, int nsweeps, int mode,
                      const double *displ, const int64_t *offset, const double *u,
                      double *E_trace, int32_t *acc_trace);
/* thermalisation helper: sweeps run with A*scale (sMC uses 2, SMC.c:110) */
int smcb_set_step_scale(smcb_engine *e, double scale);

/* ---- the host-buffer step, pipelined --------------------------------------
 * What a caller that keeps its configurations in HOST memory does per step - the reference's sMC owns R in host
 * memory and calls oneParticleMoves(R, ...) in place (SMC.c:117,195) - as ONE call: upload R (nchains*3N, AoS),
 * E <- energy + wallsEnergy (SMC.c:48), nsteps sweeps (kernel = 0, oneParticleMoves) or all-particle steps
 * (kernel = 1) from the engine's Philox stream, optionally one gather of the observables (SMC.c:137-141), download
 * the new positions into R and the chain state into E / naccept / ntrials (each nullable, nchains entries).
 * The batch is processed as four blocks of chains on four streams so one block's PCIe copies overlap the other
 * blocks' kernels; pass page-locked buffers (cudaHostAlloc / torch pin_memory) to get the overlap.  Results are
 * identical to smcb_set_positions + smcb_sweep (smcb_step_allparticle) + smcb_gather + smcb_get_positions +
 * smcb_get_chain_state. */
int smcb_sweep_host(smcb_engine *e, double *R, int nsteps, int mode, int kernel, int gather,
                    double *E, int64_t *naccept, int64_t *ntrials);

/* ---- the all-particle Smart-MC step (north-star kernel B) ----------------
 * Every particle of a chain is displaced at once, d_i = F_i A/T + xi_i, forces
 * and energy are recomputed at the proposal with a tiled O(N^2) pair kernel
 * and ONE Metropolis-Hastings test per chain decides (the reference's own
 * attempt, markovProbability, SMC.c:354-402, is dead code).
 * _fed: xi[s][c][3N] already scaled by sqrt(2A), u[s][c]; lnap (nullable)
 * receives ln(acceptance probability) [s][c]. */
int smcb_step_allparticle_fed(smcb_engine *e, int nsteps, int mode,
                              const double *xi, const double *u, double *lnap, uint8_t *accepted);
int smcb_step_allparticle(smcb_engine *e, int nsteps, int mode);

/* ---- step-size control (before production) -------------------------------
 * The reference fixes A = gamma*T for every state (main.c:48-51); a whole-configuration move (kernel = 1) with that A
 * is never accepted, and in a condensed phase the single-particle sweep's acceptance collapses too.  This runs `rounds`
 * short batches of nsteps_per_round sweeps (kernel = 0) or all-particle steps (kernel = 1) and after each multiplies
 * every chain's A by exp(gain * (its acceptance - target)) with a decreasing gain, so that A settles where the
 * acceptance is `target`.  The chains move while it runs (it is part of the thermalisation; adapting A during
 * production would break detailed balance).  Chains sharing one parameter set get their own copy.  Accept/trial
 * counters are cleared on return. */
int smcb_tune_step_size(smcb_engine *e, int kernel, int mode, double target, int rounds, int nsteps_per_round);
int smcb_get_step_sizes(smcb_engine *e, double *A);            /* nchains doubles */

/* ---- chain state -------------------------------------------------------- */
/* E: running total potential energy (set by smcb_refresh_energy or any step);
 * naccept / ntrials: accepted and attempted trials since the last reset. */
int smcb_refresh_energy(smcb_engine *e, int mode);       /* E <- energy + wallsEnergy (SMC.c:48) */
int smcb_get_chain_state(smcb_engine *e, double *E, int64_t *naccept, int64_t *ntrials);
/* caller-provided running energies, nchains doubles (sMC seeds E[0] itself, SMC.c:48, and carries
 * E[n+1] = E[n] into every sweep, SMC.c:116,194) */
int smcb_set_chain_energy(smcb_engine *e, const double *E);
int smcb_reset_counters(smcb_engine *e);

/* ---- observables (row a13 + what sMC harvests, SMC.c:137-141) ------------
 * smcb_gather adds, for every chain, into its group's block:
 *   voxel density D[33^3] and mobility Mu[33^3] (localDensityAndMobility,
 *   SMC.c:912-927; Rbin is kept per chain on the device), the z density
 *   profile (D summed over x,y), an energy histogram of E/N, and the moments
 *   n, sum E, sum E^2, sum P, sum P^2 (P = pressure + wallsPressure as the
 *   reference writes them, SMC.c:140).
 * The block is one contiguous array of uint64 counters followed by doubles per
 * group (layout from smcb_obs_layout); it is the ONLY thing ranks all-reduce.
 * Counters are exact (integer atomics) and the moments are summed over the chains of a
 * group in chain order, so the block - like the chain state - is bit-reproducible. */
typedef struct smcb_obs_layout {
    int ngroups;
    int nvox;            /* 33*33*33                                   */
    int nz;              /* 33                                         */
    int nebins;          /* energy histogram bins                      */
    double e_lo, e_hi;   /* energy-per-particle histogram range        */
    size_t u64_per_group;   /* D[nvox] Mu[nvox] zprof[nz] ehist[nebins] nsamples */
    size_t f64_per_group;   /* sumE sumE2 sumP sumP2 sumAcc            */
    size_t u64_total, f64_total;
} smcb_obs_layout;

int smcb_obs_configure(smcb_engine *e, int nebins, double e_lo, double e_hi);
int smcb_obs_layout_get(smcb_engine *e, smcb_obs_layout *out);
int smcb_gather(smcb_engine *e);
/* zero the accumulators (counters and moments).  Rbin - where every particle was at the last gather (SMC.c:921-924) -
 * is chain state and is KEPT, so a run that exports/reduces/resets after every gather counts the same mobility as
 * one that never resets.  (smcb_set_params / smcb_obs_configure start from Rbin = 0, like sMC's calloc, SMC.c:54.) */
int smcb_obs_reset(smcb_engine *e);
int smcb_obs_get(smcb_engine *e, uint64_t *counters, double *moments);           /* host copies */
/* device-side exchange for an NCCL all-reduce done by the host program: copy
 * the block to / from caller-provided DEVICE buffers (same layout). */
int smcb_obs_export_device(smcb_engine *e, void *counters_dev, void *moments_dev);
int smcb_obs_import_device(smcb_engine *e, const void *counters_dev, const void *moments_dev);
/* export + smcb_obs_reset in one ASYNCHRONOUS step on the engine's stream (smcb_stream): returns at once, so that the
 * caller's all-reduce of the exported delta - ordered after the engine's stream with an event - runs under the next
 * sweep launch instead of in front of it */
int smcb_obs_export_reset_async(smcb_engine *e, void *counters_dev, void *moments_dev);
/* One process driving several GPUs (one engine per GPU): sum the observable blocks of the n engines
 * in place with NCCL over NVLink - one ncclAllReduce(ncclUint64) of the counters and one
 * ncclAllReduce(ncclFloat64) of the moments, the only collective of the path - so that every engine
 * ends up holding the job's totals.  All engines must have the same observable layout.  NCCL is
 * loaded at run time (dlopen "libnccl.so.2"); SMCB_ERR_STATE if it is not available.  With one
 * process per GPU (torch.distributed, MPI) use smcb_obs_export_device / _import_device and the
 * launcher's own all-reduce instead.
 * The sum is IN PLACE, so it is valid ONCE per accumulation window: after it every block holds the job's totals, and
 * gathering more samples on top (or reducing again) would count the other engines' samples n times.  The engines
 * remember it: smcb_gather and a second smcb_obs_allreduce return SMCB_ERR_STATE until smcb_obs_reset.  To reduce
 * after EVERY gather, reduce deltas instead: export the block, all-reduce the copy, add it to your running total,
 * smcb_obs_reset (bench.py does this through smcb_obs_export_device). */
int smcb_obs_allreduce(smcb_engine **engines, int n);
/* release the NCCL communicators smcb_obs_allreduce caches between calls (optional) */
int smcb_obs_allreduce_teardown(void);
/* per-chain Rbin (voxel of each particle at the last gather), nchains*N ints */
int smcb_get_rbin(smcb_engine *e, int32_t *rbin);
int smcb_set_rbin(smcb_engine *e, const int32_t *rbin);

/* ---- checkpoint / resume --------------------------------------------------
 * The reference restarts from last_state_*.csv: positions only, %0.12f text (main.c:98-109,163-170;
 * the drop-in keeps that format).  The batched engine adds a binary checkpoint of EVERYTHING a
 * bit-identical continuation needs: exact positions, running energies, accept/trial counters, the
 * Philox stream identity (seed, chain0, next step), Rbin and the observable block.  A run that is
 * saved, destroyed, re-created with the same (nchains, nparticles, nsites_side), given the same
 * smcb_set_params and loaded continues exactly as the uninterrupted run would. */
int smcb_checkpoint_save(smcb_engine *e, const char *path);
int smcb_checkpoint_load(smcb_engine *e, const char *path);

/* ---- measurement -------------------------------------------------------- */
/* device time (ms, CUDA events on the engine's stream) of the kernels launched
 * by the last smcb_sweep* / smcb_step_allparticle* / smcb_evaluate call, and
 * how many kernels that call launched */
int smcb_last_kernel_ms(smcb_engine *e, float *ms, int *launches);
/* pairs inside the cutoff counted by the last sweep/step call (roofline numerator) */
int smcb_last_pair_counts(smcb_engine *e, uint64_t *pairs_total, uint64_t *pairs_in_cutoff);
/* pair distance tests the kernels of the last sweep/step call actually EXECUTED.  pairs_total above counts a sweep as the
 * reference executes it (2N(N-1) ordered pair-interactions: energySingle + forceSingle at the old and the proposed
 * position, SMC.c:300-321); the FAST kernels cache the old-position terms and screen in packed FP32, the all-particle
 * kernel visits every unordered pair once, so they execute fewer - this is the count behind roofline.frac_executed */
int smcb_last_pair_tests(smcb_engine *e, uint64_t *pair_tests_executed);
/* (debug) how the trials of the last FAST sweep launch split over the paths of k_sweep_spec: 12 counters, zero unless the
 * library was built with -DSMCB_SPEC_STATS (profiles/spec_stats.py) */
int smcb_debug_sweep_stats(smcb_engine *e, uint64_t *stats12);
/* FP64 FMA peak of this device measured with a dependent-free DFMA kernel; TFLOP/s */
int smcb_measure_fp64_peak(smcb_engine *e, double *tflops, float *ms);
/* raw device pointers for the stream-resident benchmark path (positions SoA
 * [chain][3][Npad]) */
int smcb_device_positions(smcb_engine *e, void **ptr, size_t *bytes, int *npad);
/* test hook: the FAST sweep kernel keeps per-particle energy / force / neighbour-count caches
 * (csrc/sweep_cached.cuh).  With capture on, the next FAST smcb_sweep* call leaves them here:
 * e_tot[c][N] = energySingle + wallsEnergySingle, f_tot[c][3N] = forceSingle + wallsForce,
 * nb[c][N] = partners inside the cutoff, as the kernel held them after its last trial. */
int smcb_debug_capture_cache(smcb_engine *e, int on);
int smcb_debug_get_cache(smcb_engine *e, double *e_tot, double *f_tot, double *nb);
void *smcb_stream(smcb_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
