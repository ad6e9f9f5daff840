/* oracle/shim/fftw3.h — TEST INFRASTRUCTURE (oracle build only).
 *
 * FFTW3 is the reference's one third-party dependency (SMC.h:18) and is not
 * installed here.  It is used only by fft_acf (SMC.c:1055-1090), which is off
 * the hot path.  This header declares the handful of FFTW names that function
 * uses and backs them with a naive O(n^2) DFT so that sMC() can run end to end
 * at small sizes.  <complex.h> is included before <fftw3.h> by the reference,
 * so fftw_complex is the native C99 complex double, as in real FFTW.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H
#include <complex.h>
#include <stdlib.h>
#include <math.h>

typedef double _Complex fftw_complex;
#define FFTW_ESTIMATE 64u
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

typedef struct oracle_fftw_plan_s {
    int n, kind, sign;            /* kind 0: r2c, 1: c2c */
    const double *rin;
    const fftw_complex *cin;
    fftw_complex *out;
} *fftw_plan;

static inline void *fftw_malloc(size_t n) { return malloc(n); }
static inline void fftw_free(void *p) { free(p); }

static inline fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
    (void)flags;
    fftw_plan p = malloc(sizeof(*p));
    p->n = n; p->kind = 0; p->sign = -1; p->rin = in; p->cin = NULL; p->out = out;
    return p;
}

static inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags)
{
    (void)flags;
    fftw_plan p = malloc(sizeof(*p));
    p->n = n; p->kind = 1; p->sign = sign; p->rin = NULL; p->cin = in; p->out = out;
    return p;
}

static inline void fftw_execute(const fftw_plan p)
{
    const double w0 = 2.0 * 3.14159265358979323846 / (double)p->n;
    /* NOTE: the reference passes an output array of only n/2 (+n%2) complex
     * numbers to the r2c plan (SMC.c:1067-1077) where FFTW writes n/2+1; the
     * shim writes only what the caller allocated. */
    int nout = (p->kind == 0) ? (p->n / 2 + p->n % 2) : p->n;
    for (int k = 0; k < nout; k++) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < p->n; j++) {
            double ang = p->sign * w0 * (double)(((long long)k * j) % p->n);
            double c = cos(ang), s = sin(ang);
            if (p->kind == 0) { re += p->rin[j] * c; im += p->rin[j] * s; }
            else {
                double a = creal(p->cin[j]), b = cimag(p->cin[j]);
                re += a * c - b * s; im += a * s + b * c;
            }
        }
        p->out[k] = re + im * I;
    }
}

static inline void fftw_destroy_plan(fftw_plan p) { free(p); }
#endif
