/* dropin/SMC.c — the bodies behind dropin/SMC.h: the reference's SMC.h API (SMC.c:21-1169 of
 * Kryohi/MonteCarlo-Surfacer) on top of libsmcb200.so (include/smcb200.h).
 *
 * Every physics routine here is ONE call into the CUDA engine; nothing below computes a pair
 * interaction on the host and there is no CPU fallback (a failing engine call prints
 * smcb_last_error() and aborts).  Host code that remains is what the reference also runs once per
 * run or per call, not per pair: the libc rand() stream (kept draw for draw, App. A of SURVEY.md,
 * so that `oneParticleMoves` here and in the reference agree BIT FOR BIT after the same srand),
 * the start lattice, the wall table, CSV output and the end-of-run statistics.
 *
 * Single-configuration calls (energy, forces, ..., oneParticleMoves) run on a lazily created
 * one-chain engine in STRICT arithmetic (the reference's IEEE operation order).  sMC() owns a
 * separate engine of SMCB_REPLICAS chains and advances them in lock-step with the FAST kernels
 * and the counter-based RNG, gather_lapse sweeps per launch.
 */
#include "SMC.h"
#ifndef rank
#define rank 0
#endif
#include "../../include/smcb200.h"

/* ------------------------------------------------------------------------------------ plumbing */
static smcb_engine *g_one = NULL;        /* one chain of N molecules, M*M sites */
static smcb_chain_params g_one_par;
static double g_one_W[2 * M * M];
static int g_one_have = 0;

static void smcb_die(const char *where)
{
    fprintf(stderr, "smcb200 drop-in: %s failed: %s\n", where, smcb_last_error());
    abort();
}
#define SMCB_DO(call) do { if ((call) != SMCB_OK) smcb_die(#call); } while (0)

static smcb_engine *g_pt = NULL;         /* one chain of TWO molecules: the single-point surface routines */
static smcb_chain_params g_pt_par;
static double g_pt_W[2 * M * M];
static int g_pt_have = 0;

void smcb_dropin_shutdown(void)
{
    if (g_one) smcb_destroy(g_one);
    if (g_pt) smcb_destroy(g_pt);
    g_one = g_pt = NULL;
    g_one_have = g_pt_have = 0;
}

int smcb_dropin_replicas(void)
{
    const char *s = getenv("SMCB_REPLICAS");
    const int n = s ? atoi(s) : 1;
    return n > 0 ? n : 1;
}

static int dropin_device(void)
{
    const char *s = getenv("SMCB_DEVICE");
    return s ? atoi(s) : 0;
}

static smcb_chain_params make_params(double L, double Lz, double T, double A, int wall_on)
{
    smcb_chain_params p;
    memset(&p, 0, sizeof p);
    p.L = L; p.Lz = Lz; p.T = T; p.A = A;
    p.rc2 = LJ_CUTOFF * LJ_CUTOFF;
    p.zwall_a = a0; p.zwall_b = b0;
    p.flags = wall_on ? SMCB_WALL : 0u;
    return p;
}

/* the one-chain engine with these parameters loaded (re-uploaded only when they change) */
static smcb_engine *one_chain(double L, double Lz, double T, double A, const double *W)
{
    if (!g_one) {
        SMCB_DO(smcb_create(&g_one, dropin_device(), 1, N, M));
        atexit(smcb_dropin_shutdown);
    }
    const smcb_chain_params p = make_params(L, Lz, T, A, W != NULL);
    const int same = g_one_have && memcmp(&p, &g_one_par, sizeof p) == 0 &&
                     (!W || memcmp(W, g_one_W, sizeof g_one_W) == 0);
    if (!same) {
        SMCB_DO(smcb_set_params(g_one, &p, 1, W, W ? 1 : 0, 1));
        g_one_par = p;
        if (W) memcpy(g_one_W, W, sizeof g_one_W);
        g_one_have = 1;
    }
    return g_one;
}

/* wallsEnergySingle / wallsForce take ONE point: a two-molecule chain holds it next to a dummy far outside every
 * cutoff (z is not periodic with the wall on), instead of N copies of the point in the N-molecule engine */
static smcb_engine *one_point(double rx, double ry, double rz, double L, double Lz, const double *W)
{
    if (!g_pt) {
        SMCB_DO(smcb_create(&g_pt, dropin_device(), 1, 2, M));
        if (!g_one) atexit(smcb_dropin_shutdown);
    }
    const smcb_chain_params p = make_params(L, Lz, 1.0, 1.0, 1);
    if (!(g_pt_have && memcmp(&p, &g_pt_par, sizeof p) == 0 && memcmp(W, g_pt_W, sizeof g_pt_W) == 0)) {
        SMCB_DO(smcb_set_params(g_pt, &p, 1, W, 1, 1));
        g_pt_par = p;
        memcpy(g_pt_W, W, sizeof g_pt_W);
        g_pt_have = 1;
    }
    const double r[6] = {rx, ry, rz, rx, ry, rz + 1000.0 * (L > Lz ? L : Lz)};
    SMCB_DO(smcb_set_positions(g_pt, r));
    return g_pt;
}

/* --------------------------------------------------------- LJ routines (SMC.c:557-720) ---------- */
double energySingle(const double *r, double L, int i)
{
    static double e[N];
    smcb_engine *h = one_chain(L, 1.0, 1.0, 1.0, NULL);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, e, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    return e[i];
}

/* overwrites Fx, Fy, Fz (SMC.c:589-618) */
void forceSingle(const double *r, double L, int i, double *Fx, double *Fy, double *Fz)
{
    static double f[3 * N];
    smcb_engine *h = one_chain(L, 1.0, 1.0, 1.0, NULL);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, f, NULL, NULL, NULL, NULL, NULL, NULL));
    *Fx = f[3 * i]; *Fy = f[3 * i + 1]; *Fz = f[3 * i + 2];
}

/* ACCUMULATES into F like the reference (SMC.c:656-686: F is never zeroed there) */
void forces(const double *r, double L, double *F)
{
    static double f[3 * N];
    smcb_engine *h = one_chain(L, 1.0, 1.0, 1.0, NULL);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, f, NULL, NULL, NULL, NULL, NULL, NULL));
    for (int k = 0; k < 3 * N; k++) F[k] += f[k];
}

double energy(const double *r, double L)
{
    double U = 0.0;
    smcb_engine *h = one_chain(L, 1.0, 1.0, 1.0, NULL);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, NULL, NULL, &U, NULL, NULL, NULL));
    return U;
}

double pressure(const double *r, double L, double Lz)
{
    double vir = 0.0;
    smcb_engine *h = one_chain(L, Lz, 1.0, 1.0, NULL);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, NULL, NULL, NULL, NULL, &vir, NULL));
    return -vir / (3 * L * L * Lz);
}

/* ------------------------------------------------- molecule-surface routines (SMC.c:729-895) ----- */
double wallsEnergySingle(double rx, double ry, double rz, const double *W, double L, double Lz)
{
    double e[2];
    smcb_engine *h = one_point(rx, ry, rz, L, Lz, W);
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, e, NULL, NULL, NULL, NULL, NULL));
    return e[0];
}

/* ADDS into Fx, Fy, Fz (SMC.c:773-813) */
void wallsForce(double rx, double ry, double rz, const double *W, double L, double Lz,
                double *Fx, double *Fy, double *Fz)
{
    double f[6];
    smcb_engine *h = one_point(rx, ry, rz, L, Lz, W);
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, NULL, f, NULL, NULL, NULL, NULL));
    /* the reference starts from the caller's value and adds term by term; the engine returns the
       sum formed from 0 in the same order, so add it as one term */
    *Fx += f[0]; *Fy += f[1]; *Fz += f[2];
}

double wallsEnergy(const double *r, const double *W, double L, double Lz)
{
    double U = 0.0;
    smcb_engine *h = one_chain(L, Lz, 1.0, 1.0, W);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, NULL, NULL, NULL, &U, NULL, NULL));
    return U;
}

/* the wall virial exactly as the reference writes it, quirks included (SMC.c:862-895, App. B3) */
double wallsPressure(const double *r, const double *W, double L, double Lz)
{
    double vir = 0.0;
    smcb_engine *h = one_chain(L, Lz, 1.0, 1.0, W);
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_evaluate(h, SMCB_STRICT, NULL, NULL, NULL, NULL, NULL, NULL, NULL, &vir));
    return -vir / (3 * L * L * Lz);
}

/* ----------------------------------------------------------- random inputs (matematicose.c:183) -- */
/* Gaussian vector from libc rand(): pair i draws x1 then x2 and writes A[2i] from (ln x1, cos x2),
 * A[2i+1] from (ln x2, sin x1); only 2*floor(length/2) entries are written (App. B8). */
void vecBoxMuller(double sigma, size_t length, double *A)
{
    const size_t pairs = length / 2;
    for (size_t p = 0; p < pairs; p++) {
        const double x1 = (double)rand() / (RAND_MAX + 1.0);
        const double x2 = (double)rand() / (RAND_MAX + 1.0);
        A[2 * p] = sigma * sqrt(-2 * log(1 - x1)) * cos(2 * M_PI * x2);
        A[2 * p + 1] = sigma * sqrt(-2 * log(1 - x2)) * sin(2 * M_PI * x1);
    }
}

/* ------------------------------------------------------------------- the sweep (SMC.c:278-351) -- */
/* One sweep of N sequential force-biased single-particle trials, on the GPU (STRICT kernel).  The
 * 4N+1 rand() calls of the reference's sweep are made here in its order - 3N for the Gaussian
 * displacements, one for the visiting offset, N acceptance uniforms in visiting order - and fed to
 * the engine, so after the same srand() this is the reference's trajectory bit for bit. */
void oneParticleMoves(double *R, double *Rn, const double *W, double L, double Lz, double A, double T,
                      int *j, double *U)
{
    static double displ[3 * N], u[N];
    vecBoxMuller(sqrt(2.0 * A), 3 * N, displ);
    if ((3 * N) % 2) displ[3 * N - 1] = 0.0;       /* odd length: the reference reads malloc'ed garbage here */
    const int64_t offset = rand();
    for (int n = 0; n < N; n++) u[n] = rand() / (double)RAND_MAX;

    smcb_engine *h = one_chain(L, Lz, T, A, W);
    double Etrace = 0.0;
    int32_t acc = 0;
    SMCB_DO(smcb_set_positions(h, R));
    SMCB_DO(smcb_set_chain_energy(h, U));
    /* bit-exact arithmetic at every size: one warp per chain up to N = 512, one block per chain beyond */
    SMCB_DO(smcb_sweep_traced(h, 1, SMCB_STRICT, displ, &offset, u, &Etrace, &acc));
    SMCB_DO(smcb_get_positions(h, R));
    memcpy(Rn, R, 3 * N * sizeof(double));        /* accepted: R <- Rn, rejected: Rn <- R; equal at the end */
    *j += acc;
    *U = Etrace;
}

/* ------------------------------------------------------------ start state (SMC.c:413-501) ------- */
/* fcc lattice of Na x Na x Nz cells shifted by a/4 and wrapped with periods (L, L, Lz - Lz/20).
 * For n = 4 Na^2 Nz with Na = floor(cbrt(n/4)) (108, 256, 500 ...) this is the reference's lattice
 * (its random jitter integer-divides to zero, App. B7).  For other n the reference writes an
 * invalid lattice (e.g. 4096: 96 coincident molecules); here Na grows until the cells tile n. */
void initializeBox(double L, double Lz, int n, double *X)
{
    srand(42);
    const int cells = n / 4;
    int Na = 1;
    while ((Na + 1) * (Na + 1) * (Na + 1) <= cells) Na++;
    while (Na < cells && cells % (Na * Na) != 0) Na++;
    if (cells % (Na * Na) != 0 || 4 * cells != n)
        perror("Can't make a crystal with this N, it should be an integer times a perfect square, all divisible by 4.\n");
    const int Nz = cells / (Na * Na) > 0 ? cells / (Na * Na) : 1;
    const double a = L / Na, Pz = Lz - Lz / 20.0;
    static const double basis[4][3] = {{0, 0, 0}, {.5, .5, 0}, {.5, 0, .5}, {0, .5, .5}};
    int p = 0;
    for (int i = 0; i < Na; i++)
        for (int jy = 0; jy < Na; jy++)
            for (int k = 0; k < Nz; k++)
                for (int b = 0; b < 4 && p < n; b++, p++) {
                    X[3 * p] = a * i + a * basis[b][0];
                    X[3 * p + 1] = a * jy + a * basis[b][1];
                    X[3 * p + 2] = a * k + a * basis[b][2];
                }
    for (int q = 0; q < n; q++)
        for (int c = 0; c < 3; c++) {
            (void)rand();                                  /* the reference draws (and discards) a jitter here */
            X[3 * q + c] += a / 4;
        }
    for (int q = 0; q < n; q++) {
        X[3 * q] -= L * rint(X[3 * q] / L);
        X[3 * q + 1] -= L * rint(X[3 * q + 1] / L);
        X[3 * q + 2] -= Pz * rint(X[3 * q + 2] / Pz);
    }
    if (n == N && boundsCheck(X, L, Lz - 0.5) > 0)
        perror("Lz is too small or there is something else going wrong\n");
}

/* per-site (a, b) = (x0^12, x0^6) * ymin with x0 ~ N(x0m, x0sigma), ymin ~ N(ym, ymsigma), written
 * to W interleaved, m = i*M + j, and to the wall CSV (which is closed here, as in the reference) */
void initializeWalls(double x0m, double x0sigma, double ymm, double ymsigma, double *W, FILE *wall)
{
    srand(42);
    double *X0 = calloc(M * M, sizeof(double));     /* calloc: for odd M*M the last entry is never drawn */
    double *YM = calloc(M * M, sizeof(double));
    vecBoxMuller(x0sigma, M * M, X0);
    vecBoxMuller(ymsigma, M * M, YM);
    if (wall) fprintf(wall, "nx, ny, x0, ymin\n");
    for (int m = 0; m < M * M; m++) {
        const double x0 = X0[m] + x0m, ymin = YM[m] + ymm;
        if (wall) fprintf(wall, "%d, %d, %f, %f\n", m / M, m % M, x0, ymin);
        W[2 * m] = pow(x0, 12.0) * ymin;
        W[2 * m + 1] = pow(x0, 6.) * ymin;
    }
    free(X0); free(YM);
    if (wall) fclose(wall);
}

void shiftSystem(double *r, double L)
{
    for (int k = 0; k < 3 * N; k++) r[k] -= L * rint(r[k] / L);
}

void shiftSystem2D(double *r, double L)
{
    for (int n = 0; n < N; n++) {
        r[3 * n] -= L * rint(r[3 * n] / L);
        r[3 * n + 1] -= L * rint(r[3 * n + 1] / L);
    }
}

void shiftSystem3D(double *r, double L, double Lz)
{
    shiftSystem2D(r, L);
    for (int n = 0; n < N; n++) r[3 * n + 2] -= Lz * rint(r[3 * n + 2] / Lz);
}

/* molecules outside the x,y box are counted and returned; molecules beyond a wall only reported */
int boundsCheck(double *r, double L, double Lz)
{
    int escaped = 0;
    for (int n = 0; n < N; n++) {
        if (fabs(r[3 * n]) > L / 2.0 || fabs(r[3 * n + 1]) > L / 2.0) {
            printf("Particles are escaping the system and going to the beta-carotene Valhalla\n");
            escaped++;
        } else if (fabs(r[3 * n + 2]) > Lz / 2.0) {
            printf("Particles are smashing the walls :(\n");
        }
    }
    return escaped;
}

/* ---------------------------------------------------------------- observables (SMC.c:912-964) --- */
/* cumulative 33^3 voxel counts D, per-molecule voxel Rbin and mobility counts Mu, on the GPU */
void localDensityAndMobility(const double *r, double L, double Lz, unsigned long int *D, int *Rbin,
                             unsigned long int *Mu)
{
    smcb_engine *h = one_chain(L, Lz, 1.0, 1.0, NULL);
    smcb_obs_layout lay;
    SMCB_DO(smcb_obs_reset(h));
    SMCB_DO(smcb_obs_layout_get(h, &lay));
    uint64_t *cnt = malloc(lay.u64_total * sizeof(uint64_t));
    int32_t rb[N];
    for (int n = 0; n < N; n++) rb[n] = Rbin[n];
    SMCB_DO(smcb_set_positions(h, r));
    SMCB_DO(smcb_set_rbin(h, rb));
    SMCB_DO(smcb_gather(h));
    SMCB_DO(smcb_obs_get(h, cnt, NULL));
    SMCB_DO(smcb_get_rbin(h, rb));
    for (int v = 0; v < lay.nvox; v++) { D[v] += cnt[v]; Mu[v] += cnt[lay.nvox + v]; }
    for (int n = 0; n < N; n++) Rbin[n] = rb[n];
    free(cnt);
}

/* z edges of the non-uniform grid: LAYER_DEPTH-thick layers at both walls, the rest in the middle */
void createZRange(double Lz, double *z_cells)
{
    const int half = (Ncz - 2) / 2;
    for (int k = 0; k < half; k++) {
        z_cells[k] = LAYER_DEPTH * k;
        z_cells[Ncz - 1 - k] = Lz - LAYER_DEPTH * k;
    }
    const double middle = Lz - (Ncz - 4) * LAYER_DEPTH;
    z_cells[half] = Lz / 2 - middle / 6;
    z_cells[Ncz / 2] = Lz / 2 + middle / 6;
}

/* The non-uniform variant is dead code in the reference (its only call is commented out,
 * SMC.c:29-30) and is outside the GPU path: host loop, same binning rule with z_cells edges. */
void localDensityAndMobility_nonuniz(const double *r, double L, double Lz, double *z_cells,
                                     unsigned long int *D, int *Rbin, unsigned long int *Mu)
{
    for (int n = 0; n < N; n++) {
        const int i = (int)floor((r[3 * n] / L + .5) * Ncx), jy = (int)floor((r[3 * n + 1] / L + .5) * Ncx);
        const double z = r[3 * n + 2] + Lz / 2;
        int k = 0;
        while (k + 1 < Ncz && z >= z_cells[k + 1]) k++;
        const int v = i * Ncx * Ncz + jy * Ncz + k;
        D[v]++;
        if (Rbin[n] != v) { Mu[v]++; Rbin[n] = v; }
    }
}

/* Common-neighbour analysis AS THE REFERENCE COMPUTES IT (SMC.c:971-1045), so that LCA[] is the array the
 * reference would fill, entry for entry.  For every pair l > i it records: bonded (distance below LCA_cutoff,
 * minimum image in x,y), how many particles q < l are bonded to both, and how many consecutive ones of those are
 * bonded to each other.  The reference addresses pair (l,i) at slot (l-1)(l-2)/2 + i, which makes the LAST slot of
 * row l also the FIRST slot of row l+1, never clears a slot, looks the (q,i) bond up at (q-1)(q-2)/2 + i even when
 * i > q, and keeps appending to one 8-entry scratch list per slot; all of that decides what ends up in LCA[], so it
 * is kept: slot arithmetic below is the reference's, written once in `slot_of`.  (Host-side and off the GPU path:
 * sMC samples it every LCA_TIME gathers, and SURVEY App. B6 explains why the numbers sMC derives from it are of
 * little use.)  The scratch list is bounded here; the reference overruns its own beyond 8 shared neighbours. */
static inline long slot_of(long hi, long lo) { return (hi * hi - 3 * hi + 2) / 2 + lo; }

void clusterAnalysis(const double *r, int N_, double L, int *LCA)
{
    const long nslots = ((long)N_ * N_ - N_) / 2;
    unsigned char *bonded = calloc((size_t)nslots, 1);
    int *shared = calloc((size_t)nslots, sizeof(int));
    int *linked = calloc((size_t)nslots, sizeof(int));
    int list[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 1; l < N_; l++)
        for (int i = 0; i < l; i++) {
            double dx = r[3 * l] - r[3 * i];
            dx = dx - L * rint(dx / L);
            double dy = r[3 * l + 1] - r[3 * i + 1];
            dy = dy - L * rint(dy / L);
            const double dz = r[3 * l + 2] - r[3 * i + 2];
            if (dx * dx + dy * dy + dz * dz < LCA_cutoff * LCA_cutoff) bonded[slot_of(l, i)] = 1;
        }
    for (int l = 1; l < N_; l++)
        for (int i = 0; i < l; i++) {
            const long s = slot_of(l, i);
            if (!bonded[s]) continue;
            for (int q = 0; q < l; q++) {
                if (q == i) continue;
                if (bonded[slot_of(l, q)] & bonded[slot_of(q, i)]) {
                    if (shared[s] < 8) list[shared[s]] = q;
                    shared[s]++;
                }
            }
            for (int m = 1; m < shared[s] && m < 8; m++)
                if (bonded[slot_of(list[m], list[m - 1])]) linked[s]++;
        }
    for (long n = 0; n < nslots; n++) {
        if (shared[n] > 6) printf("LCA cutoff might be too big, clustering data will be corrupted\n");
        LCA[3 * n] = bonded[n]; LCA[3 * n + 1] = shared[n]; LCA[3 * n + 2] = linked[n];
    }
    free(bonded); free(shared); free(linked);
}

/* ------------------------------------------------------- autocorrelation (SMC.c:1055-1142) ------- */
/* The reference's fft_acf hands FFTW (an un-vendored dependency, version unpinned) a real-to-complex plan of the
 * mean-free series, keeps the first lfft = length/2 (+1 if odd) bins F_j, and takes a COMPLEX backward transform of
 * length lfft of |F_j|^2 (SMC.c:1067-1085):  acf[k] = Re sum_{j<lfft} |F_j|^2 e^{+2 pi i jk/lfft} / sum_j |F_j|^2.
 * (With half the spectrum on half the length, entry k is the circular autocorrelation at lag 2k up to an end-bin
 * term - a quirk of the reference that plotting.jl and tau = sum(acf) inherit, kept.)  FFTW's arithmetic is
 * restated here by its definition, the discrete Fourier transform, evaluated in O(n log n) for ANY length with
 * Bluestein's chirp-z identity over a power-of-two FFT: the reference's 16e6-sweep series are transformed whole,
 * not truncated.  tests/test_postproc_parity.py pins it against the reference's fft_acf to 1e-10. */
typedef struct { double re, im; } cplx;

static void fft_pow2(cplx *a, size_t n, int sign)            /* in place, n a power of two, e^{sign 2 pi i jk/n} */
{
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { const cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1;
        const double w0 = sign * 6.283185307179586476925286766559 / (double)len;
        for (size_t k = 0; k < half; k++) {
            const double wr = cos(w0 * (double)k), wi = sin(w0 * (double)k);
            for (size_t s0 = k; s0 < n; s0 += len) {
                cplx *u = a + s0, *v = a + s0 + half;
                const double xr = v->re * wr - v->im * wi, xi = v->re * wi + v->im * wr;
                v->re = u->re - xr; v->im = u->im - xi;
                u->re += xr; u->im += xi;
            }
        }
    }
}

/* out[k] = sum_j in[j] e^{sign 2 pi i jk/n}, k < nout, for any n: jk = (j^2 + k^2 - (k-j)^2)/2 turns the sum into a
 * convolution with the chirp c_m = e^{sign pi i m^2/n} (phases from m^2 mod 2n, exact in integers) */
static int dft_any(const cplx *in, size_t n, int sign, cplx *out, size_t nout)
{
    if (n == 0) return 0;
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1;
    cplx *chirp = malloc(n * sizeof(cplx)), *x = calloc(m, sizeof(cplx)), *y = calloc(m, sizeof(cplx));
    if (!chirp || !x || !y) { free(chirp); free(x); free(y); return -1; }
    const unsigned long long two_n = 2ull * n;
    for (size_t j = 0; j < n; j++) {
        const unsigned long long q = ((unsigned long long)j * j) % two_n;
        const double ang = sign * 3.1415926535897932384626433832795 * (double)q / (double)n;
        chirp[j].re = cos(ang); chirp[j].im = sin(ang);
    }
    for (size_t j = 0; j < n; j++) {                         /* x_j = in_j c_j ;  y_j = conj(c_j), wrapped */
        x[j].re = in[j].re * chirp[j].re - in[j].im * chirp[j].im;
        x[j].im = in[j].re * chirp[j].im + in[j].im * chirp[j].re;
        y[j].re = chirp[j].re; y[j].im = -chirp[j].im;
        if (j) y[m - j] = y[j];
    }
    fft_pow2(x, m, -1);
    fft_pow2(y, m, -1);
    for (size_t j = 0; j < m; j++) {
        const double pr = x[j].re * y[j].re - x[j].im * y[j].im, pi = x[j].re * y[j].im + x[j].im * y[j].re;
        x[j].re = pr; x[j].im = pi;
    }
    fft_pow2(x, m, +1);
    for (size_t k = 0; k < nout && k < n; k++) {             /* out_k = c_k (x * y)_k / m */
        const double vr = x[k].re / (double)m, vi = x[k].im / (double)m;
        out[k].re = vr * chirp[k].re - vi * chirp[k].im;
        out[k].im = vr * chirp[k].im + vi * chirp[k].re;
    }
    free(chirp); free(x); free(y);
    return 0;
}

DoubleArray fft_acf(const double *H, size_t length, int k_max)
{
    DoubleArray acf;
    if (length < (size_t)k_max * 2 + 1) {
        k_max = (int)rint((double)(length / 2)) - 2;
        printf("Number of datapoints too low to calculate autocorrelation, new k_max: %d\n", k_max);
    }
    if (k_max < 1) k_max = 1;
    acf.length = (size_t)k_max;
    acf.data = calloc((size_t)k_max, sizeof(double));
    const size_t lfft = length / 2 + length % 2;
    cplx *z = malloc((length ? length : 1) * sizeof(cplx)), *f = malloc((lfft ? lfft : 1) * sizeof(cplx)), *c = malloc((lfft ? lfft : 1) * sizeof(cplx));
    if (!acf.data || !z || !f || !c || lfft == 0) { free(z); free(f); free(c); return acf; }
    const double meanH = mean(H, length);
    for (size_t i = 0; i < length; i++) { z[i].re = H[i] - meanH; z[i].im = 0.0; }
    int bad = dft_any(z, length, -1, f, lfft);               /* the bins the r2c plan writes into lfft slots */
    for (size_t j = 0; j < lfft; j++) { f[j].re = f[j].re * f[j].re + f[j].im * f[j].im; f[j].im = 0.0; }
    if (!bad) bad = dft_any(f, lfft, +1, c, lfft);           /* FFTW_BACKWARD, unnormalised */
    if (!bad)
        for (int k = 0; k < k_max && (size_t)k < lfft; k++) acf.data[k] = c[k].re / c[0].re;
    else
        fprintf(stderr, "fft_acf: out of memory for a series of %zu points\n", length);
    free(z); free(f); free(c);
    return acf;
}

/* the reference's direct-sum variant (SMC.c:1094-1121), kept for its callers: sums over i < length-k_max-1 for
 * every lag, normalised by the lag-0 value */
void simple_acf(const double *H, size_t length, int k_max, double *acf)
{
    if (length < (size_t)k_max * 2) perror("error: number of datapoints too low to calculate autocorrelation");
    const double m = mean(H, length);
    const long upto = (long)length - k_max - 1;
    for (int k = 0; k < k_max; k++) {
        double c = 0.0;
        for (long i = 0; i < upto; i++) c += (H[i] - m) * (H[i + k] - m);
        acf[k] = c / (double)((long)length - k_max);
    }
    for (int k = k_max - 1; k >= 0; k--) acf[k] = acf[k] / acf[0];
}

/* variance of every tau-th sample (decorrelated subsample) */
double variance_corr(const double *A, double tau, size_t length)
{
    int stride = (int)floor(tau);
    if (stride < 1) stride = 1;
    const size_t m = length / (size_t)stride;
    if (m < 1000) printf("\nThere doesn't seem to be enough data to compute the variance\n");
    const double mu = mean(A, length);
    double acc = 0.0;
    for (size_t k = 0; k < m; k++) acc += (A[k * stride] - mu) * (A[k * stride] - mu);
    return m > 1 ? acc / (double)(m - 1) : 0.0;
}

/* ------------------------------------------------------------- the simulation (SMC.c:21-267) ---- */
static FILE *open_csv(const char *kind, double rho, double T, const char *header)
{
    char name[96];
    snprintf(name, sizeof name, "./%s_N%d_M%d_r%0.4f_T%0.2f_rank%d.csv", kind, N, M, rho, T, rank);
    FILE *f = fopen(name, "w");
    if (!f) perror("error while opening csv files");
    else if (header) fprintf(f, "%s", header);
    return f;
}

static void dump_voxels(FILE *f, const uint64_t *D, const uint64_t *Mu, const uint64_t *D_old, const uint64_t *Mu_old)
{
    if (!f) return;
    for (int i = 0; i < Ncx; i++)
        for (int jy = 0; jy < Ncx; jy++)
            for (int k = 0; k < Ncz; k++) {
                const int v = i * Ncx * Ncz + jy * Ncz + k;
                fprintf(f, "%d, %d, %d, %lu, %lu\n", i, jy, k, (unsigned long)(D[v] - (D_old ? D_old[v] : 0)),
                        (unsigned long)(Mu[v] - (Mu_old ? Mu_old[v] : 0)));
            }
}

/* sweeps in bounded launches; appends sMC's per-sweep records of replica 0 */
static void run_sweeps(smcb_engine *h, int C, int nsweeps, double *E_out, int *acc_out)
{
    enum { CHUNK = 4096 };
    double *Et = malloc((size_t)CHUNK * C * sizeof(double));
    int32_t *At = malloc((size_t)CHUNK * C * sizeof(int32_t));
    for (int done = 0; done < nsweeps;) {
        const int n = nsweeps - done < CHUNK ? nsweeps - done : CHUNK;
        SMCB_DO(smcb_sweep_traced(h, n, SMCB_FAST, NULL, NULL, NULL, Et, At));
        for (int s = 0; s < n; s++) { E_out[done + s] = Et[(size_t)s * C]; acc_out[done + s] = At[(size_t)s * C]; }
        done += n;
    }
    free(Et); free(At);
}

struct Sim sMC(double L, double Lz, double T, double A, const double *W, const double *R0,
               int maxsteps, int gather_lapse, int eqsteps)
{
    const double rho = N / (L * L * Lz);
    const int gather_steps = maxsteps / gather_lapse;
    const int C = smcb_dropin_replicas();
    const int Nc = Ncx * Ncx * Ncz;
    clock_t start, end;

    /* stream identity: like srand(time(NULL)) (SMC.c:40), or SMCB_SEED for a reproducible run */
    const char *seed_env = getenv("SMCB_SEED");
    const uint64_t seed = seed_env ? strtoull(seed_env, NULL, 10) : ((uint64_t)time(NULL) << 8) ^ (uint64_t)rank;

    smcb_engine *h = NULL;
    SMCB_DO(smcb_create(&h, dropin_device(), C, N, M));
    const smcb_chain_params par = make_params(L, Lz, T, A, 1);
    SMCB_DO(smcb_set_params(h, &par, 1, W, 1, 1));
    SMCB_DO(smcb_broadcast_positions(h, R0));
    SMCB_DO(smcb_set_rng(h, seed, 0, 0));
    SMCB_DO(smcb_refresh_energy(h, SMCB_FAST));            /* E[0] = energy + wallsEnergy (SMC.c:48) */

    double *E = calloc((size_t)maxsteps + 1, sizeof(double));
    double *Eeq = calloc((size_t)(eqsteps > 0 ? eqsteps : 1), sizeof(double));
    double *P = calloc((size_t)(gather_steps > 0 ? gather_steps : 1), sizeof(double));
    int *jj = calloc((size_t)(maxsteps > 0 ? maxsteps : 1), sizeof(int));
    int *jt = calloc((size_t)(eqsteps > 0 ? eqsteps : 1), sizeof(int));
    double *Ec = malloc((size_t)C * sizeof(double));
    double *R = malloc((size_t)C * 3 * N * sizeof(double));
    smcb_obs_layout lay;
    SMCB_DO(smcb_obs_layout_get(h, &lay));
    uint64_t *cnt = calloc(lay.u64_total, sizeof(uint64_t));
    uint64_t *cnt_old = calloc(lay.u64_total, sizeof(uint64_t));
    double l2[7] = {0}, l3[7] = {0}, l1 = 0.0;
    int *clusters = calloc((size_t)3 * N * (N - 1) / 2, sizeof(int));

    FILE *positions = open_csv("positions", rho, T, NULL);
    if (positions) {
        for (int n = 0; n < N; n++) fprintf(positions, "x%d,y%d,z%d,", n + 1, n + 1, n + 1);
        fprintf(positions, "\n");
        for (int k = 0; k < 3 * N; k++) fprintf(positions, "%0.3lf,", R0[k]);
        fprintf(positions, "\n");
    }
    FILE *data = open_csv("data", rho, T, "E, P, jj\n");
    FILE *local = open_csv("local", rho, T, "nx, ny, nz, n, mu\n");
    FILE *local_temp = open_csv("local_temp", rho, T, "nx, ny, nz, n, mu\n");
    FILE *total_clusters = open_csv("total_clusters", rho, T, "l1, l2, l3\n");
    FILE *autocorrelation = open_csv("autocorrelation", rho, T, "CH\n");

    printf("\nStarting new run with %d particles in %0.1fx%0.1fx%0.1f box, ", N, L, L, Lz);
    printf("T=%0.2f, rho=%0.4f, A=%0.3f, for %d steps (%d replica chain%s on the GPU)...\n", T, rho, A, maxsteps, C, C > 1 ? "s" : "");

    /* ---- thermalisation with A*2 (SMC.c:110-125) ---- */
    SMCB_DO(smcb_get_chain_state(h, Ec, NULL, NULL));
    const double E_start = Ec[0];
    start = clock();
    SMCB_DO(smcb_set_step_scale(h, 2.0));
    run_sweeps(h, C, eqsteps, Eeq, jt);
    SMCB_DO(smcb_set_step_scale(h, 1.0));
    end = clock();
    double sim_time = ((double)(end - start)) / CLOCKS_PER_SEC;
    int *now = currentTime();
    printf("\nThermalization completed in %0.1f mins at %02d:%02d, with ", sim_time / 60, now[0], now[1]);
    if (eqsteps > 0)
        printf("average acceptance ratio %0.3f, mean energy %0.3f.\n", intmean(jt, eqsteps) / N, mean(Eeq, eqsteps) + 3 * N * T / 2);
    else
        printf("no thermalisation sweeps.\n");

    /* ---- production (SMC.c:134-196): gather k happens before sweep n = k*gather_lapse - 1 ---- */
    if (eqsteps > 0) printf("The expected time of execution is ~%0.1f minutes.\n", 1.03 * sim_time * maxsteps / eqsteps / 60);
    start = clock();
    SMCB_DO(smcb_get_chain_state(h, Ec, NULL, NULL));
    /* E[0] is the energy of the thermalised state (the reference keeps the stale pre-thermalisation
       value here, App. B10, which shifts its whole E[] series by a constant) */
    E[0] = Ec[0];
    (void)E_start;
    SMCB_DO(smcb_obs_reset(h));
    int n = 0;
    for (int k = 1; k <= gather_steps; k++) {
        const int upto = k * gather_lapse - 1;             /* sweeps completed before this gather */
        run_sweeps(h, C, upto - n, E + 1 + n, jj + n);
        n = upto;
        double vir_lj[1], vir_wall[1];
        /* P = pressure + wallsPressure (SMC.c:140), stored at k-1 (the reference writes P[k], one past
           the end of its array and leaves P[0] = 0, App. B4) */
        if (C == 1) {
            SMCB_DO(smcb_evaluate(h, SMCB_FAST, NULL, NULL, NULL, NULL, NULL, NULL, vir_lj, vir_wall));
        } else {
            double *vl = malloc((size_t)2 * C * sizeof(double));
            SMCB_DO(smcb_evaluate(h, SMCB_FAST, NULL, NULL, NULL, NULL, NULL, NULL, vl, vl + C));
            vir_lj[0] = vl[0]; vir_wall[0] = vl[C];
            free(vl);
        }
        P[k - 1] = -vir_lj[0] / (3 * L * L * Lz) + -vir_wall[0] / (3 * L * L * Lz);
        SMCB_DO(smcb_gather(h));                           /* localDensityAndMobility of every replica (SMC.c:141) */
        if (k % LCA_TIME == 0 || k % STORAGE_TIME == 0) SMCB_DO(smcb_get_positions(h, R));
        if (k % LCA_TIME == 0 && gather_steps >= LCA_TIME) {
            const double w = 1.0 / (double)(gather_steps / LCA_TIME);
            clusterAnalysis(R, N, L, clusters);
            for (int i = 0; i < N * (N - 1) / 2; i++)
                if (clusters[3 * i]) {
                    l1 += w;
                    if (clusters[3 * i + 1] < 7) l2[clusters[3 * i + 1]] += w;
                    if (clusters[3 * i + 2] < 7) l3[clusters[3 * i + 2]] += w;
                }
        }
        if (k % STORAGE_TIME == 0) {
            if (positions) {
                for (int i = 0; i < 3 * N; i++) fprintf(positions, "%0.3lf,", R[i]);
                fprintf(positions, "\n");
            }
            printf("\rStoring the latest density distribution at %d steps... ", n + 1);
            SMCB_DO(smcb_obs_get(h, cnt, NULL));
            dump_voxels(local_temp, cnt, cnt + Nc, cnt_old, cnt_old + Nc);
            memcpy(cnt_old, cnt, lay.u64_total * sizeof(uint64_t));
        }
    }
    run_sweeps(h, C, maxsteps - n, E + 1 + n, jj + n);
    end = clock();
    sim_time = ((double)(end - start)) / CLOCKS_PER_SEC;
    printf("\n\nTime: %0.1f s (%0.1f per million)\n", sim_time, sim_time * 1e6 / (maxsteps > 0 ? maxsteps : 1));

    /* ---- data preparation and storage (SMC.c:203-240) ---- */
    for (int k = 0; k < gather_steps; k++) P[k] += rho * T;
    for (int k = 0; k <= maxsteps; k++) E[k] += 3 * N * T / 2;
    if (data)
        for (int k = 0; k < gather_steps; k++) fprintf(data, "%0.9lf, %0.9lf, %d\n", E[(size_t)k * gather_lapse], P[k], jj[k]);
    SMCB_DO(smcb_obs_get(h, cnt, NULL));
    dump_voxels(local, cnt, cnt + Nc, NULL, NULL);
    /* the reference opens total_clusters_*.csv, writes its header and reports the numbers on stdout only (SMC.c:92, 226-231) */
    printf("l1[1] = %0.9f\n", l1);
    printf("l2[0] = %0.9f\tl2[1] = %0.9f\tl2[2] = %0.9f\tl2[3] = %0.9f\tl2[4] = %0.9f\tl2[5] = %0.9f\n", l2[0], l2[1], l2[2], l2[3], l2[4], l2[5]);
    printf("l3[0] = %0.9f\tl3[1] = %0.9f\tl3[2] = %0.9f\tl3[3] = %0.9f\tl3[4] = %0.9f\tl3[5] = %0.9f\n", l3[0], l3[1], l3[2], l3[3], l3[4], l3[5]);

    DoubleArray acf = fft_acf(E, (size_t)maxsteps + 1, KMAX);
    const double tau = sum(acf.data, acf.length);
    if (autocorrelation)
        for (size_t m = 0; m < acf.length; m++) fprintf(autocorrelation, "%0.6lf\n", acf.data[m]);

    Sim results;
    memset(&results, 0, sizeof results);
    results.E = mean(E, (size_t)maxsteps + 1);
    results.dE = sqrt(fmax(0.0, variance(E, (size_t)maxsteps + 1)));
    results.P = gather_steps > 0 ? mean(P, gather_steps) : 0.0;
    results.dP = gather_steps > 0 ? sqrt(fmax(0.0, variance(P, gather_steps))) : 0.0;
    results.acceptance_ratio = maxsteps > 0 ? intmean(jj, maxsteps) / N : 0.0;
    results.tau = tau;
    results.cv = variance(E, (size_t)maxsteps + 1) / (T * T);
    SMCB_DO(smcb_get_positions(h, R));
    memcpy(results.Rfinal, R, 3 * N * sizeof(double));
    for (int s = 0; s < 7; s++) { results.l2[s] = l2[s]; results.l3[s] = l3[s]; }
    results.ACF = acf;

    free(E); free(Eeq); free(P); free(jj); free(jt); free(Ec); free(R); free(cnt); free(cnt_old); free(clusters);
    if (positions) fclose(positions);
    if (data) fclose(data);
    if (local) fclose(local);
    if (local_temp) fclose(local_temp);
    if (total_clusters) fclose(total_clusters);
    if (autocorrelation) fclose(autocorrelation);
    SMCB_DO(smcb_destroy(h));
    return results;
}
