/* dropin/matematicose.c — host-side scalar helpers with the reference's names and meaning
 * (matematicose.c:8-278).  They are off the GPU path: sMC uses sum/mean/variance/intmean for its
 * end-of-run statistics (SMC.c:244-250) and vecBoxMuller (which lives in SMC.c here, next to the
 * rand() stream it consumes).  Where the reference has a defect the fixed behaviour is noted. */
#ifndef SMCB_DROPIN_MATEMATICOSE_C
#define SMCB_DROPIN_MATEMATICOSE_C
#include <stdio.h>
#include <stdlib.h>
#include "matematicose.h"

/* |a-b| < 1e-12, the tolerance the north star also uses for fp64 parity (matematicose.c:8-14) */
bool isPicoEqual(double a, double b) { return fabs(a - b) < 1e-12; }

/* declared by the reference header but never defined there; relative 1e-9 */
bool isApproxEqual(double a, double b) { return fabs(a - b) <= 1e-9 * fmax(1.0, fmax(fabs(a), fabs(b))); }

double sum(const double *A, size_t length)
{
    double acc = 0.0;
    for (size_t k = 0; k < length; k++) acc += A[k];
    return acc;
}

int intsum(const int *A, size_t length)
{
    int acc = 0;
    for (size_t k = 0; k < length; k++) acc += A[k];
    return acc;
}

double mean(const double *A, size_t length) { return sum(A, length) / (double)length; }

/* integer accumulator like the reference (matematicose.c:53-60), then a double division */
double intmean(const int *A, size_t length) { return (double)intsum(A, length) / (double)length; }

/* population variance <A^2> - <A>^2 (matematicose.c:95-102), without the temporary array */
double variance(const double *A, size_t length)
{
    double s1 = 0.0, s2 = 0.0;
    for (size_t k = 0; k < length; k++) { s1 += A[k]; s2 += A[k] * A[k]; }
    const double m = s1 / (double)length;
    return s2 / (double)length - m * m;
}

/* the reference writes A[length] and skips A[0] (App. B9); this zeroes A[0..length-1] */
void zeros(size_t length, double *A)
{
    for (size_t k = 0; k < length; k++) A[k] = 0.0;
}

void elforel(const double *A, const double *B, double *C, size_t length)
{
    for (size_t k = 0; k < length; k++) C[k] = A[k] * B[k];
}

void pointwise(double (*f)(double), double *A, size_t length)
{
    for (size_t k = 0; k < length; k++) A[k] = f(A[k]);
}

int double_max_index(double *A, size_t length)
{
    size_t best = 0;
    for (size_t k = 1; k < length; k++) if (A[k] > A[best]) best = k;
    return (int)best;
}

int double_min_index(double *A, size_t length)
{
    size_t best = 0;
    for (size_t k = 1; k < length; k++) if (A[k] < A[best]) best = k;
    return (int)best;
}

/* ---- root finding: secant iteration on g(x) = f(x) - c until inf < g < sup ----------------------- */
static double secant_core(double (*f)(double), double c, double xa, double xb, double inf, double sup)
{
    double ga = f(xa) - c, gb = f(xb) - c;
    if (ga > inf && ga < sup) return xa;
    if (gb > inf && gb < sup) return xb;
    if (ga * gb > 0) {
        perror("f(X1) and f(X2) must have an opposing sign");
        return -1;
    }
    for (int it = 0; it < 10000 && !(gb > inf && gb < sup); it++) {
        const double xn = xb - gb * (xb - xa) / (gb - ga);
        xa = xb; ga = gb;
        xb = xn; gb = f(xb) - c;
    }
    return xb;
}

double zerosecant(double (*f)(double), double x1, double x2, double inf, double sup)
{
    return secant_core(f, 0.0, x1, x2, inf, sup);
}

double secant(double (*f)(double), double c, double x1, double x2, double inf, double sup)
{
    return secant_core(f, c, x1, x2, inf, sup);
}

/* last sign change of f - c on [x1, x2], scanned downwards in 1000 steps, refined by the secant */
double findzero_last(double (*f)(double), double c, double x1, double x2, double inf, double sup)
{
    const double h = (x2 - x1) / 1000;
    for (int k = 0; k < 1000; k++) {
        const double hi = x2 - k * h, lo = hi - h;
        if ((f(hi) - c) * (f(lo) - c) < 0) return secant_core(f, c, lo, hi, inf, sup);
    }
    perror("no zeros found");
    return -1;
}

/* upward recurrence J[l+1] = (2l+1)/x J[l] - J[l-1]; J[0], J[1] given */
void fast_bessel(double x, double lmax, double *J)
{
    for (int l = 1; l < lmax; l++) J[l + 1] = ((2 * l + 1) / x) * J[l] - J[l - 1];
}

/* ---- finite differences and quadrature ------------------------------------------------------------ */
double der3(double *F, int x, double h) { return (F[x + 1] - F[x - 1]) / (2 * h); }

double der5(double *F, int x, double h)
{
    return (8 * (F[x + 1] - F[x - 1]) - (F[x + 2] - F[x - 2])) / (12 * h);
}

double der5_c(double (*f)(double), double x, double h)
{
    return (8 * (f(x + h) - f(x - h)) - (f(x + 2 * h) - f(x - 2 * h))) / (12 * h);
}

/* composite Simpson over samples 0..xmax-1 spaced h (pairs of intervals) */
double simpson_integral(double *fun, int xmax, double h)
{
    double acc = 0.0;
    for (int k = 1; k + 1 < xmax; k += 2) acc += fun[k - 1] + 4.0 * fun[k] + fun[k + 1];
    return acc * h / 3.0;
}

/* ---- 1-D gradient descent with the 5-point derivative ----------------------------------------------- */
static double descend(double (*f)(double), double x, double rate, double h)
{
    for (int it = 0; it < 1000000; it++) {
        const double g = der5_c(f, x, h);
        if (fabs(g) <= 1e-7) break;
        x -= rate * g;
    }
    return x;
}

double grad_descent_1D(double (*f)(double), double x1, double x2)
{
    const double start = (x2 - x1) / 2;
    const double rate = fabs(f(x2) - f(start)) / 200;
    return descend(f, start, rate, (x2 - x1) / 5e4);
}

/* 64 starting points, best minimum wins.  (The reference's starting points all collapse onto x1
 * because rand()/RAND_MAX is an integer division, matematicose.c:262; here they are spread.) */
double stochastic_grad_descent_1D(double (*f)(double), double x1, double x2)
{
    enum { STARTS = 64 };
    double x[STARTS], fx[STARTS];
    srand(42);
    for (int k = 0; k < STARTS; k++) x[k] = x1 + (x2 - x1) * ((double)rand() / ((double)RAND_MAX + 1.0));
    const double rate = fabs(f(x2) - f(x[STARTS / 2 - 1])) / 200;
    for (int k = 0; k < STARTS; k++) {
        x[k] = descend(f, x[k], rate, (x2 - x1) / 5e4);
        fx[k] = f(x[k]);
    }
    return fx[double_min_index(fx, STARTS)];
}
#endif
