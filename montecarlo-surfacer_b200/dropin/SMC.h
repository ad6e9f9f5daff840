/* dropin/SMC.h — the reference's SMC.h API (SMC.h:25-121 of Kryohi/MonteCarlo-Surfacer) re-hosted on
 * libsmcb200.so, the B200 engine.  A driver written against the reference - its own main.c does
 * `#include "SMC.c"` (main.c:2) - compiles UNCHANGED with `-I<this directory>` and links
 * `-lsmcb200`; see INTEGRATION.md.
 *
 * What is kept, because callers see it: the size/physics macros (SMC.h:26-61), the result records
 * `DoubleArray` and `Sim` field for field (SMC.h:71-88; sMC returns Sim by value and main.c frees
 * ACF.data, main.c:173), every prototype of SMC.h:92-121, the single-translation-unit habit of
 * pulling matematicose.c and misccose.c in from here (SMC.h:19-20) and the caller-owned AoS buffers.
 * What changed: the bodies (SMC.c in this directory) call the CUDA engine; <fftw3.h> is not needed.
 * N and M stay compile-time macros for source compatibility, but may now be given with -DN= -DM=. */
#ifndef SMCB_DROPIN_SMC_H
#define SMCB_DROPIN_SMC_H

#include <errno.h>
#include <limits.h>
#include <math.h>
#include <signal.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

#include "matematicose.c"
#include "misccose.c"

/* ---- sizes (SMC.h:26-29) ------------------------------------------------------------------- */
#ifndef M
#define M 3            /* surface sites per side, M*M in total */
#endif
#ifndef N
#define N 108          /* molecules */
#endif

/* ---- physics constants (SMC.h:32-39) --------------------------------------------------------- */
#define a0 5.960464477539063e-9     /* flat wall, repulsive coefficient  */
#define b0 2.44140625e-5            /* flat wall, attractive coefficient */
#define TRUNCATE 1
#if TRUNCATE == 1
#define LJ_CUTOFF 3.0
#else
#define LJ_CUTOFF (L / 2)
#endif

/* ---- harvesting cadence and grids (SMC.h:43-61) ---------------------------------------------- */
#define STORAGE_TIME 1000
#define LCA_TIME 10
#define LCA_cutoff 1.7
#define Ncx 33
#define Ncz 33
#define LAYER_DEPTH 5.0
#define KMAX 2500000

/* ---- result records (SMC.h:71-88) ------------------------------------------------------------ */
typedef struct DoubleArray {
    size_t length;
    double *data;
} DoubleArray;

typedef struct Sim {
    double E, dE;                 /* mean total energy and its standard deviation       */
    double P, dP;                 /* mean pressure and its standard deviation           */
    double acceptance_ratio;
    double cv;                    /* var(E)/T^2                                         */
    double tau;                   /* integrated autocorrelation time of E               */
    double Rfinal[3 * N];
    double l2[7], l3[7];          /* common-neighbour statistics                        */
    struct DoubleArray ACF;       /* data is heap memory owned by the caller            */
} Sim;

/* ---- the API (SMC.h:92-121) -------------------------------------------------------------------- */
struct Sim sMC(double L, double Lz, double T, double A, const double *W, const double *R0,
               int maxsteps, int gather_lapse, int eqsteps);
void vecBoxMuller(double sigma, size_t length, double *A);
void shiftSystem(double *r, double L);
void shiftSystem2D(double *r, double L);
void shiftSystem3D(double *r, double L, double Lz);
void createZRange(double Lz, double *z_cells);
void initializeWalls(double x0m, double x0sigma, double ym, double ymsigma, double *W, FILE *wall);
void initializeBox(double L, double Lz, int n, double *X);

void oneParticleMoves(double *R, double *Rn, const double *W, double L, double Lz, double A, double T,
                      int *j, double *U);

double energySingle(const double *r, double L, int i);
void forceSingle(const double *r, double L, int i, double *Fx, double *Fy, double *Fz);
void forces(const double *r, double L, double *F);
double energy(const double *r, double L);
double pressure(const double *r, double L, double Lz);
double wallsEnergy(const double *r, const double *W, double L, double Lz);
double wallsEnergySingle(double rx, double ry, double rz, const double *W, double L, double Lz);
void wallsForce(double rx, double ry, double rz, const double *W, double L, double Lz,
                double *Fx, double *Fy, double *Fz);
double wallsPressure(const double *r, const double *W, double L, double Lz);
void localDensityAndMobility(const double *r, double L, double Lz, unsigned long int *D, int *Rbin,
                             unsigned long int *Mu);
void localDensityAndMobility_nonuniz(const double *r, double L, double Lz, double *z_cells,
                                     unsigned long int *D, int *Rbin, unsigned long int *Mu);
void clusterAnalysis(const double *r, int N_, double L, int *LCA);
int boundsCheck(double *r, double L, double Lz);

void simple_acf(const double *H, size_t length, int k_max, double *acf);
DoubleArray fft_acf(const double *H, size_t length, int k_max);
double variance_corr(const double *A, double tau, size_t length);

/* ---- additions of the drop-in (not in the reference) ------------------------------------------- */
/* Number of independent replica chains sMC advances in lock-step on the GPU (environment variable
 * SMCB_REPLICAS, default 1).  Replica 0 is the chain the returned Sim describes; the voxel
 * histograms written to local_*.csv are summed over replicas. */
int smcb_dropin_replicas(void);
/* release the engine handles the wrappers created lazily */
void smcb_dropin_shutdown(void);

#endif /* SMCB_DROPIN_SMC_H */
