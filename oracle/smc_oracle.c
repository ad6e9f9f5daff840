/* oracle/smc_oracle.c — CPU restatement of the reference's Smart-Monte-Carlo path.
 *
 * *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py may load this.  The product
 * (libsmcb200.so) never links, imports or calls anything under oracle/.
 *
 * What it is: the algorithm of /root/reference/SMC.c (+ matematicose.c's
 * Box-Muller) restated with RUNTIME sizes (the reference fixes N and M with
 * macros, SMC.h:26-29) and with the random numbers passed in explicitly, so a
 * test can drive this, the compiled reference (oracle/_ref/libref_*.so) and the
 * CUDA path with one stream.  Every routine keeps the reference's IEEE operation
 * ORDER (no FMA: build with -ffp-contract=off) so it is bit-identical to the
 * reference built the same way; tests/test_oracle_vs_ref.py pins that.
 *
 * Parity status: PINNED against the reference itself, compiled here from its own
 * sources by oracle/build_ref.sh (the reference has no tests or golden vectors
 * of its own), and against tests/golden/ fixtures generated from
 * that build (tests/golden/make_golden.py).
 *
 * Two additions have no reference counterpart and say so where they are defined:
 * the bulk switch (`periodic_z`, following SMC_noMPI_noWall.c's 3-D minimum
 * image) and the all-particle step (`orc_allparticle_step`), whose only
 * reference is the commented-out markovProbability (SMC.c:354-402).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct orc_sys {
    int N;            /* particles                     (SMC.h:29  #define N) */
    int M;            /* wall sites per side, M*M      (SMC.h:26  #define M) */
    double L;         /* x,y period                    (main.c:35-44)        */
    double Lz;        /* wall-to-wall distance; z period in bulk mode        */
    double rc2;       /* LJ cutoff squared             (SMC.h:36 LJ_CUTOFF=3) */
    double a0, b0;    /* flat z-wall 12-6 coefficients (SMC.h:32-33)         */
    int periodic_z;   /* 1 = bulk: z gets the minimum image too (SMC_noMPI_noWall.c:470-475) */
    int wall;         /* 1 = molecule-surface potential on (SMC.c:729-813)   */
} orc_sys;

size_t orc_sizeof_sys(void) { return sizeof(orc_sys); }

/* ---------------------------------------------------------------- geometry */

/* d - P*rint(d/P), the reference's minimum-image idiom (SMC.c:568, 602, ...) */
static inline double min_image(double d, double P) { return d - P * rint(d / P); }

/* displacement a-b under the box rules; returns |d|^2 summed x,y,z left to right */
static inline double pair_sep(const orc_sys *s, const double *a, const double *b, double d[3])
{
    d[0] = min_image(a[0] - b[0], s->L);
    d[1] = min_image(a[1] - b[1], s->L);
    d[2] = a[2] - b[2];                       /* z: no wrap in slab mode (SMC.c:571-572) */
    if (s->periodic_z) d[2] = min_image(d[2], s->Lz);
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

/* ------------------------------------------------------- LJ, one particle  */

/* SMC.c:557-583.  4 * sum_{l != i, r2 < rc2} ( r^-12 - r^-6 ), l ascending. */
double orc_energy_single(const orc_sys *s, const double *r, int i)
{
    double acc = 0.0, d[3];
    for (int l = 0; l < s->N; l++) {
        if (l == i) continue;
        double r2 = pair_sep(s, r + 3 * l, r + 3 * i, d);
        if (r2 < s->rc2) {
            double r6 = r2 * r2 * r2;
            acc += 1.0 / (r6 * r6) - 1.0 / r6;
        }
    }
    return acc * 4;
}

/* SMC.c:589-618.  F_i = sum (48 r^-14 - 24 r^-8) (r_i - r_l); overwrites F[0..2]. */
void orc_force_single(const orc_sys *s, const double *r, int i, double *F)
{
    double fx = 0.0, fy = 0.0, fz = 0.0, d[3];
    for (int l = 0; l < s->N; l++) {
        if (l == i) continue;
        double r2 = pair_sep(s, r + 3 * i, r + 3 * l, d);
        if (r2 < s->rc2) {
            double r8 = r2 * r2 * r2 * r2;
            double g = 48.0 / (r8 * r2 * r2 * r2) - 24.0 / r8;
            fx += g * d[0];
            fy += g * d[1];
            fz += g * d[2];
        }
    }
    F[0] = fx; F[1] = fy; F[2] = fz;
}

/* ------------------------------------------------------ LJ, whole system   */

/* SMC.c:626-646 (note the r^-12 is built as a 6-fold product of r2 here). */
double orc_energy(const orc_sys *s, const double *r)
{
    double acc = 0.0, d[3];
    for (int l = 1; l < s->N; l++)
        for (int i = 0; i < l; i++) {
            double r2 = pair_sep(s, r + 3 * l, r + 3 * i, d);
            if (r2 < s->rc2)
                acc += 1.0 / (r2 * r2 * r2 * r2 * r2 * r2) - 1.0 / (r2 * r2 * r2);
        }
    return acc * 4;
}

/* SMC.c:656-686.  Newton-3 all pairs; ACCUMULATES into F (caller zeroes). */
void orc_forces(const orc_sys *s, const double *r, double *F)
{
    double d[3];
    for (int l = 1; l < s->N; l++)
        for (int i = 0; i < l; i++) {
            double r2 = pair_sep(s, r + 3 * l, r + 3 * i, d);
            if (r2 < s->rc2) {
                double r8 = r2 * r2 * r2 * r2;
                double g = 24.0 / r8 - 48.0 / (r8 * r2 * r2 * r2);
                for (int c = 0; c < 3; c++) {
                    F[3 * l + c] -= g * d[c];
                    F[3 * i + c] += g * d[c];
                }
            }
        }
}

/* SMC.c:696-720.  LJ virial pressure; volume L*L*Lz. */
double orc_pressure(const orc_sys *s, const double *r)
{
    double acc = 0.0, d[3];
    for (int l = 1; l < s->N; l++)
        for (int i = 0; i < l; i++) {
            double r2 = pair_sep(s, r + 3 * l, r + 3 * i, d);
            if (r2 < s->rc2) {
                double r6 = r2 * r2 * r2;
                acc += 24.0 / r6 - 48.0 / (r6 * r6);
            }
        }
    return -acc / (3 * s->L * s->L * s->Lz);
}

/* ------------------------------------------------------------- the surface */

/* signed distance to the nearer of the two walls at z = -/+ Lz/2, with the
 * reference's clamp for particles outside the slab (SMC.c:735-739, 783-786). */
static inline double wall_dz(double rz, double Lz)
{
    double dz = rz + Lz / 2;
    dz = dz - Lz * rint(dz / Lz);
    if (rz <= -Lz / 2.0) dz = 0.0001;
    else if (rz >= Lz / 2) dz = -0.0001;
    return dz;
}

/* SMC.c:729-763 */
double orc_walls_energy_single(const orc_sys *s, double rx, double ry, double rz, const double *W)
{
    if (!s->wall) return 0.0;
    const int M = s->M;
    double acc = 0.0;
    double dw = s->L / M;
    double dz = wall_dz(rz, s->Lz);
    double z6 = dz * dz * dz * dz * dz * dz;
    acc += s->a0 / (z6 * z6) - s->b0 / z6;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            int m = j + i * M;
            double dx = min_image(rx - i * dw, s->L);
            double dy = min_image(ry - j * dw, s->L);
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < s->rc2) {
                double r6 = r2 * r2 * r2;
                acc += W[2 * m] / (r6 * r6) - W[2 * m + 1] / r6;
            }
        }
    return acc * 4;
}

/* SMC.c:773-813.  ADDS into F[0..2]. */
void orc_walls_force(const orc_sys *s, double rx, double ry, double rz, const double *W, double *F)
{
    if (!s->wall) return;
    const int M = s->M;
    double dw = s->L / M;
    double dz = wall_dz(rz, s->Lz);
    double z8 = dz * dz * dz * dz * dz * dz * dz * dz;
    double g = 48.0 * s->a0 / (z8 * dz * dz * dz * dz * dz * dz) - 24.0 * s->b0 / z8;
    F[2] += g * dz;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            int m = j + i * M;
            double dx = min_image(rx - i * dw, s->L);
            double dy = min_image(ry - j * dw, s->L);
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < s->rc2) {
                double r8 = r2 * r2 * r2 * r2;
                g = 48.0 * W[2 * m] / (r8 * r2 * r2 * r2) - 24.0 * W[2 * m + 1] / r8;
                F[0] += g * dx;
                F[1] += g * dy;
                F[2] += g * dz;
            }
        }
}

/* SMC.c:822-859.  One running sum over all particles (not a sum of a4's). */
double orc_walls_energy(const orc_sys *s, const double *r, const double *W)
{
    if (!s->wall) return 0.0;
    const int M = s->M;
    double acc = 0.0;
    double dw = s->L / M;
    for (int n = 0; n < s->N; n++) {
        double dz = wall_dz(r[3 * n + 2], s->Lz);
        double z6 = dz * dz * dz * dz * dz * dz;
        acc += s->a0 / (z6 * z6) - s->b0 / z6;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) {
                int m = j + i * M;
                double dx = min_image(r[3 * n] - i * dw, s->L);
                double dy = min_image(r[3 * n + 1] - j * dw, s->L);
                double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < s->rc2) {
                    double r6 = r2 * r2 * r2;
                    acc += W[2 * m] / (r6 * r6) - W[2 * m + 1] / r6;
                }
            }
    }
    return acc * 4;
}

/* SMC.c:862-895, AS WRITTEN: dz uses L/2 where Lz/2 was
 * meant, there is no clamp, and the flat-wall term is added once per in-cutoff
 * site.  Site-major loop order. */
double orc_walls_pressure(const orc_sys *s, const double *r, const double *W)
{
    if (!s->wall) return 0.0;
    const int M = s->M;
    double acc = 0.0;
    double dw = s->L / M;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            int m = j + i * M;
            for (int n = 0; n < s->N; n++) {
                double dx = min_image(r[3 * n] - i * dw, s->L);
                double dy = min_image(r[3 * n + 1] - j * dw, s->L);
                double dz = r[3 * n + 2] + s->L / 2;
                dz = dz - s->Lz * rint(dz / s->Lz);
                double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < s->rc2) {
                    double r6 = r2 * r2 * r2;
                    acc += 24.0 * W[2 * m + 1] / r6 - 48.0 * W[2 * m] / (r6 * r6);
                    double z6 = dz * dz * dz * dz * dz * dz;
                    acc += 24.0 * s->b0 / z6 - 48.0 * s->a0 / (z6 * z6);
                }
            }
        }
    return -acc / (3 * s->L * s->L * s->Lz);
}

/* The wall virial sum the reference MEANT (no reference code: wallsPressure has the three defects listed above):
 * per particle, distance to the nearer wall as in wallsEnergySingle (SMC.c:735-739, clamp included), the flat-wall
 * term r dV/dr once, and every surface site inside the cutoff.  Returns the SUM (pressure = -sum / (3 L^2 Lz)). */
double orc_walls_virial_intended(const orc_sys *s, const double *r, const double *W)
{
    if (!s->wall) return 0.0;
    const int M = s->M;
    const double dw = s->L / M;
    double acc = 0.0;
    for (int n = 0; n < s->N; n++) {
        const double rz = r[3 * n + 2];
        double dz = rz + s->Lz / 2;
        dz = dz - s->Lz * rint(dz / s->Lz);
        if (rz <= -s->Lz / 2.0) dz = 0.0001;
        else if (rz >= s->Lz / 2) dz = -0.0001;
        const double z6 = dz * dz * dz * dz * dz * dz;
        acc += 24.0 * s->b0 / z6 - 48.0 * s->a0 / (z6 * z6);
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) {
                const int m = j + i * M;
                const double dx = min_image(r[3 * n] - i * dw, s->L);
                const double dy = min_image(r[3 * n + 1] - j * dw, s->L);
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < s->rc2) {
                    const double r6 = r2 * r2 * r2;
                    acc += 24.0 * W[2 * m + 1] / r6 - 48.0 * W[2 * m] / (r6 * r6);
                }
            }
    }
    return acc;
}

/* ------------------------------------------------------------ random input */

/* matematicose.c:183-193.  `rnd` holds the rand() results in draw order; uses
 * 2*floor(len/2) of them, fills the same number of outputs (an odd tail entry
 * is left untouched, as in the reference).  Note the crossed pairing. */
void orc_box_muller(double sigma, size_t len, const int *rnd, int rand_max, double *out)
{
    const double two_pi = 2 * M_PI;
    for (size_t i = 0; i < len / 2; i++) {
        double x1 = (double)rnd[2 * i] / (rand_max + 1.0);
        double x2 = (double)rnd[2 * i + 1] / (rand_max + 1.0);
        out[2 * i] = sigma * sqrt(-2 * log(1 - x1)) * cos(two_pi * x2);
        out[2 * i + 1] = sigma * sqrt(-2 * log(1 - x2)) * sin(two_pi * x1);
    }
}

/* ------------------------------------------------------- the sweep (a1)    */

/* SMC.c:278-351 with its random inputs made explicit:
 *   displ[3N]  the vecBoxMuller(sqrt(2A),3N) output      (SMC.c:284)
 *   offset     the rand() that picks the first particle  (SMC.c:290)
 *   u[N]       rand()/(double)RAND_MAX per trial, in VISITING order (SMC.c:335)
 * R is updated in place, Rn is the scratch copy, *naccept += accepted trials,
 * *E += sum of accepted (Un-Um).  `accepted` (nullable) gets one flag per trial
 * in visiting order. */
void orc_sweep(const orc_sys *s, double *R, double *Rn, const double *W, double A, double T,
               const double *displ, long long offset, const double *u,
               int *naccept, double *E, unsigned char *accepted)
{
    const int N = s->N;
    memcpy(Rn, R, 3 * (size_t)N * sizeof(double));
    for (int nn = 0; nn < N; nn++) {
        int n = (int)((nn + offset) % N);      /* 64-bit: the reference's int sum can overflow (App. A) */
        double *p = R + 3 * n, *q = Rn + 3 * n;
        double Fm[3], Fn[3], dl[3];

        double Um = orc_energy_single(s, R, n) + orc_walls_energy_single(s, p[0], p[1], p[2], W);
        orc_force_single(s, R, n, Fm);
        orc_walls_force(s, p[0], p[1], p[2], W, Fm);

        for (int c = 0; c < 3; c++) {
            dl[c] = Fm[c] * A / T + displ[3 * n + c];
            q[c] = p[c] + dl[c];
        }
        q[0] = q[0] - s->L * rint(q[0] / s->L);         /* SMC.c:315-316: x,y only */
        q[1] = q[1] - s->L * rint(q[1] / s->L);
        if (s->periodic_z) q[2] = q[2] - s->Lz * rint(q[2] / s->Lz);   /* bulk extension */

        double Un = orc_energy_single(s, Rn, n) + orc_walls_energy_single(s, q[0], q[1], q[2], W);
        orc_force_single(s, Rn, n, Fn);
        orc_walls_force(s, q[0], q[1], q[2], W, Fn);

        double gx = Fn[0] - Fm[0], gy = Fn[1] - Fm[1], gz = Fn[2] - Fm[2];
        double dW = (gx * gx + gy * gy + gz * gz + 2.0 * (gx * Fm[0] + gy * Fm[1] + gz * Fm[2])) * A / (4.0 * T);
        double ap = exp(-(Un - Um + (dl[0] * (Fn[0] + Fm[0]) + dl[1] * (Fn[1] + Fm[1]) + dl[2] * (Fn[2] + Fm[2])) / 2.0 + dW) / T);

        int ok = u[nn] < ap;
        if (ok) {
            p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
            *naccept += 1;
            *E += Un - Um;
        } else {
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
        }
        if (accepted) accepted[nn] = (unsigned char)ok;
    }
}

/* The same sweep driven by the raw integer stream the reference would have
 * pulled from rand(): 3N ints for Box-Muller, 1 for the offset, N for the
 * trials.  rnd must hold 4N+1 values (N even). */
void orc_sweep_from_ints(const orc_sys *s, double *R, double *Rn, const double *W, double A, double T,
                         const int *rnd, int rand_max, int *naccept, double *E)
{
    const int N = s->N;
    double *displ = calloc(3 * (size_t)N, sizeof(double));
    double *u = malloc((size_t)N * sizeof(double));
    orc_box_muller(sqrt(2.0 * A), 3 * (size_t)N, rnd, rand_max, displ);
    long long offset = rnd[2 * (3 * (size_t)N / 2)];
    const int *tail = rnd + 2 * (3 * (size_t)N / 2) + 1;
    for (int nn = 0; nn < N; nn++) u[nn] = (double)tail[nn] / (double)rand_max;
    orc_sweep(s, R, Rn, W, A, T, displ, offset, u, naccept, E, NULL);
    free(displ); free(u);
}

/* host-side expansion of the same integer stream into (displ, offset, u), for
 * feeding the CUDA path the numbers the reference consumed */
void orc_expand_stream(int N, double A, const int *rnd, int rand_max,
                       double *displ, long long *offset, double *u)
{
    memset(displ, 0, 3 * (size_t)N * sizeof(double));
    orc_box_muller(sqrt(2.0 * A), 3 * (size_t)N, rnd, rand_max, displ);
    *offset = rnd[2 * (3 * (size_t)N / 2)];
    const int *tail = rnd + 2 * (3 * (size_t)N / 2) + 1;
    for (int nn = 0; nn < N; nn++) u[nn] = (double)tail[nn] / (double)rand_max;
}

/* ------------------------------------------------------------ observables  */

/* SMC.c:912-927.  Cumulative voxel counts; bin indices pass through uint8_t
 * exactly as in the reference. */
void orc_local_density(const orc_sys *s, const double *r, int ncx, int ncz,
                       unsigned long *D, int *Rbin, unsigned long *Mu)
{
    for (int n = 0; n < s->N; n++) {
        uint8_t i = floor((r[3 * n] / s->L + .5) * ncx);
        uint8_t j = floor((r[3 * n + 1] / s->L + .5) * ncx);
        uint8_t k = floor((r[3 * n + 2] / s->Lz + .5) * ncz);
        int v = i * ncx * ncz + j * ncz + k;
        D[v]++;
        if (Rbin[n] != v) { Mu[v]++; Rbin[n] = v; }
    }
}

/* SMC.c:529-543 without the printf: particles outside |x|,|y| <= L/2 */
int orc_bounds_check(const orc_sys *s, const double *r, double Lz_check, int *beyond_wall)
{
    int out = 0, bw = 0;
    for (int j = 0; j < s->N; j++) {
        if (fabs(r[3 * j]) > s->L / 2.0 || fabs(r[3 * j + 1]) > s->L / 2.0) out++;
        else if (fabs(r[3 * j + 2]) > Lz_check / 2.0) bw++;
    }
    if (beyond_wall) *beyond_wall = bw;
    return out;
}

/* ------------------------------------------------------------- start state */

/* SMC.c:413-465 as it actually behaves: fcc cells Na x Na x Nz, every site
 * shifted by a/4 (the jitter term integer-divides to 0, SMC.c:456-458), then
 * wrapped with periods (L, L, Lz - Lz/20).  Returns the number of sites written
 * (== n only when n = 4*Na*Na*Nz; the reference is invalid otherwise, App. B7). */
int orc_initialize_box(double L, double Lz, int n, double *X)
{
    int Nc = (int)ceil(n / 4);
    int Na = 1;
    for (int nc = 1; nc < n; nc++)
        if (nc * nc * nc > Nc) { Na = nc - 1; break; }
    int Nz = (int)rint((n / 4) / (Na * Na));
    double a = L / Na;
    int sites = 0;
    for (int i = 0; i < Na; i++)
        for (int j = 0; j < Na; j++)
            for (int k = 0; k < Nz; k++) {
                int c = i * Na * Nz + j * Nz + k;
                if (4 * (c + 1) > n) continue;            /* never write past 3n doubles */
                double *x = X + 12 * c;
                x[0] = a * i;          x[1] = a * j;          x[2] = a * k;
                x[3] = a * i + a / 2;  x[4] = a * j + a / 2;  x[5] = a * k;
                x[6] = a * i + a / 2;  x[7] = a * j;          x[8] = a * k + a / 2;
                x[9] = a * i;          x[10] = a * j + a / 2; x[11] = a * k + a / 2;
                sites += 4;
            }
    double Pz = Lz - Lz / 20.0;
    for (int p = 0; p < n; p++) {
        for (int c = 0; c < 3; c++) X[3 * p + c] += a / 4 + L * 0 / 50;
        X[3 * p] = X[3 * p] - L * rint(X[3 * p] / L);
        X[3 * p + 1] = X[3 * p + 1] - L * rint(X[3 * p + 1] / L);
        X[3 * p + 2] = X[3 * p + 2] - Pz * rint(X[3 * p + 2] / Pz);
    }
    return sites;
}

/* Corrected generator for any n = 4*nx*ny*nz (needed for N=4096, where the
 * reference's cell count is wrong): same fcc basis and a/4 shift, cell edge
 * L/nx in all three directions, wrapped like the reference. */
void orc_fcc_lattice(double L, double Lz, int nx, int ny, int nz, double *X)
{
    double a = L / nx;
    static const double basis[4][3] = {{0, 0, 0}, {.5, .5, 0}, {.5, 0, .5}, {0, .5, .5}};
    double Pz = Lz - Lz / 20.0;
    size_t p = 0;
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++)
                for (int b = 0; b < 4; b++, p++) {
                    double x = a * i + a * basis[b][0] + a / 4;
                    double y = a * j + a * basis[b][1] + a / 4;
                    double z = a * k + a * basis[b][2] + a / 4;
                    X[3 * p] = x - L * rint(x / L);
                    X[3 * p + 1] = y - L * rint(y / L);
                    X[3 * p + 2] = z - Pz * rint(z / Pz);
                }
}

/* SMC.c:475-501 given the two Gaussian vectors it draws (X0 ~ N(0,x0sigma),
 * YM ~ N(0,ymsigma), each M*M long): a = x0^12 * ymin, b = x0^6 * ymin. */
void orc_walls_from_gauss(int M, double x0m, double ymm, const double *X0, const double *YM, double *W)
{
    for (int m = 0; m < M * M; m++) {
        double x0 = X0[m] + x0m;
        W[2 * m] = pow(x0, 12.0) * (YM[m] + ymm);
        W[2 * m + 1] = pow(x0, 6.) * (YM[m] + ymm);
    }
}

/* ------------------------------------------- counter-based RNG (new build) */

/* Philox4x32-10 (Salmon et al., SC'11), the generator the CUDA path uses for
 * its production streams; restated here so CPU and GPU can replay one stream.
 * No reference counterpart: the reference uses libc rand(). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in (0,1): ((hi:lo) >> 11) + 0.5) * 2^-53 */
static inline double u53(uint32_t lo, uint32_t hi)
{
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

/* Stream layout shared with the CUDA path (csrc/philox.cuh):
 *   key     = (seed_lo, seed_hi ^ 0x5MCB...) -- see smcb_rng_key
 *   counter = (step_lo, step_hi, chain, particle | tag<<28)
 * tag 0/1: the two Philox blocks that give a particle's 3 Gaussians,
 * tag 2: the trial uniform (kernel A) ; particle=0,tag 3: per-step scalars
 * (sweep offset / whole-chain uniform).
 * Outputs 3 STANDARD normals (sigma applied by the caller) and one uniform. */
void orc_rng_particle(uint64_t seed, uint32_t chain, uint64_t step, uint32_t particle,
                      double g[3], double *u)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), chain, particle};
    uint32_t a[4], b[4];
    orc_philox4x32_10(ctr, key, a);
    ctr[3] = particle | (1u << 28);
    orc_philox4x32_10(ctr, key, b);
    double u1 = u53(a[0], a[1]), u2 = u53(a[2], a[3]);
    double u3 = u53(b[0], b[1]), u4 = u53(b[2], b[3]);
    double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
    g[0] = r1 * cos(2 * M_PI * u2);
    g[1] = r1 * sin(2 * M_PI * u2);
    g[2] = r2 * cos(2 * M_PI * u4);
    if (u) {
        uint32_t c[4];
        ctr[3] = particle | (2u << 28);
        orc_philox4x32_10(ctr, key, c);
        *u = u53(c[0], c[1]);
    }
}

void orc_rng_step_scalars(uint64_t seed, uint32_t chain, uint64_t step, uint32_t *offset, double *u)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), chain, 3u << 28};
    uint32_t c[4];
    orc_philox4x32_10(ctr, key, c);
    if (offset) *offset = c[2] & 0x7fffffffu;
    if (u) *u = u53(c[0], c[1]);
}

/* ------------------------------------- whole-configuration force / energy  */

/* Per-particle total force (LJ + surface) and the configuration's energies,
 * assembled from the reference's own routines: F_i = forceSingle(i) followed by
 * wallsForce (SMC.c:303-304), U_lj = energy(), U_wall = wallsEnergy(),
 * virial = the pair sum inside pressure() before the -1/(3 L^2 Lz) factor. */
void orc_total(const orc_sys *s, const double *r, const double *W,
               double *F, double *U_lj, double *U_wall, double *virial)
{
    for (int i = 0; i < s->N; i++) {
        orc_force_single(s, r, i, F + 3 * i);
        orc_walls_force(s, r[3 * i], r[3 * i + 1], r[3 * i + 2], W, F + 3 * i);
    }
    if (U_lj) *U_lj = orc_energy(s, r);
    if (U_wall) *U_wall = orc_walls_energy(s, r, W);
    if (virial) *virial = orc_pressure(s, r) * (-(3 * s->L * s->L * s->Lz));
}

/* -------------------------------------------- all-particle SMC step (B)    */

/* One whole-configuration Smart-MC trial.  NO WORKING REFERENCE EXISTS: the
 * reference's attempt (markovProbability, SMC.c:354-402) is commented out as
 * "doesn't work" and computes a per-component acceptance.  This is the same
 * Rossky-Doll-Friedman move as the live single-particle one (SMC.c:307-329)
 * applied to every particle at once, with the ONE scalar Metropolis-Hastings
 * test the summed proposal densities give (SURVEY.md §8a derivation):
 *     d_i  = F_i * A/T + xi_i                       xi ~ N(0, 2A)
 *     ln a = -[ U' - U + 1/2 sum d_i.(F'_i + F_i) + A/(4T) sum (|F'_i|^2 - |F_i|^2) ] / T
 * accept iff u < exp(ln a).  R, F, U are the current state (F, U must be
 * consistent with R on entry) and are replaced on acceptance.
 * xi[3N] are the displacement noises already scaled by sqrt(2A).
 * Returns 1 if accepted; *lnap_out gets ln a. */
int orc_allparticle_step(const orc_sys *s, double *R, double *F, double *U, const double *W,
                         double A, double T, const double *xi, double u, double *lnap_out)
{
    const int N = s->N;
    double *Rn = malloc(3 * (size_t)N * sizeof(double));
    double *Fn = malloc(3 * (size_t)N * sizeof(double));
    double *dl = malloc(3 * (size_t)N * sizeof(double));
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 3; c++) {
            dl[3 * i + c] = F[3 * i + c] * A / T + xi[3 * i + c];
            Rn[3 * i + c] = R[3 * i + c] + dl[3 * i + c];
        }
        Rn[3 * i] = Rn[3 * i] - s->L * rint(Rn[3 * i] / s->L);
        Rn[3 * i + 1] = Rn[3 * i + 1] - s->L * rint(Rn[3 * i + 1] / s->L);
        if (s->periodic_z) Rn[3 * i + 2] = Rn[3 * i + 2] - s->Lz * rint(Rn[3 * i + 2] / s->Lz);
    }
    double Ulj, Uw;
    orc_total(s, Rn, W, Fn, &Ulj, &Uw, NULL);
    double Un = Ulj + Uw;
    double drift = 0.0, f2 = 0.0;
    for (int i = 0; i < 3 * N; i++) {
        drift += dl[i] * (Fn[i] + F[i]);
        f2 += Fn[i] * Fn[i] - F[i] * F[i];
    }
    double lnap = -((Un - *U) + drift / 2.0 + f2 * A / (4.0 * T)) / T;
    if (lnap_out) *lnap_out = lnap;
    int ok = u < exp(lnap);
    if (ok) {
        memcpy(R, Rn, 3 * (size_t)N * sizeof(double));
        memcpy(F, Fn, 3 * (size_t)N * sizeof(double));
        *U = Un;
    }
    free(Rn); free(Fn); free(dl);
    return ok;
}

/* ------------------------------------------------------ timing helper      */

/* cpu_baseline leg of bench.py: `nsweeps` reference-semantics sweeps of one
 * chain driven by a cheap LCG-fed integer stream (the timing must not depend
 * on glibc's rand lock).  Returns accepted trials. */
long orc_run_sweeps(const orc_sys *s, double *R, const double *W, double A, double T,
                    int nsweeps, unsigned seed, double *E)
{
    const int N = s->N;
    double *Rn = malloc(3 * (size_t)N * sizeof(double));
    int *rnd = malloc((4 * (size_t)N + 1) * sizeof(int));
    uint64_t st = 0x9E3779B97F4A7C15ull ^ seed;
    long acc = 0;
    for (int k = 0; k < nsweeps; k++) {
        for (size_t i = 0; i < 4 * (size_t)N + 1; i++) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            rnd[i] = (int)((st >> 33) & 0x7fffffff);
        }
        int j = 0;
        orc_sweep_from_ints(s, R, Rn, W, A, T, rnd, 2147483647, &j, E);
        acc += j;
    }
    free(Rn); free(rnd);
    return acc;
}
