/* dropin/matematicose.h — prototypes of the reference's scalar helpers (matematicose.h:6-28).
 * Host-side C; only the statistics (sum/mean/variance/intmean) are used by sMC (SMC.c:244-250). */
#ifndef SMCB_DROPIN_MATEMATICOSE_H
#define SMCB_DROPIN_MATEMATICOSE_H
#include <math.h>
#include <stdbool.h>
#include <stddef.h>

bool isPicoEqual(double a, double b);
bool isApproxEqual(double a, double b);
void pointwise(double (*f)(double), double *A, size_t length);
int double_max_index(double *A, size_t length);
int double_min_index(double *A, size_t length);
double sum(const double *A, size_t length);
int intsum(const int *A, size_t length);
double mean(const double *A, size_t length);
double intmean(const int *A, size_t length);
double variance(const double *A, size_t length);
double variance_corr(const double *A, double tau, size_t length);
void zeros(size_t length, double *A);
void elforel(const double *A, const double *B, double *C, size_t length);
double zerosecant(double (*f)(double), double x1, double x2, double inf, double sup);
double secant(double (*f)(double), double c, double x1, double x2, double inf, double sup);
double findzero_last(double (*f)(double), double c, double x1, double x2, double inf, double sup);
void fast_bessel(double x, double lmax, double *J);
double der3(double *F, int x, double h);
double der5(double *F, int x, double h);
double der5_c(double (*f)(double), double x, double h);
double simpson_integral(double *fun, int xmax, double h);
double grad_descent_1D(double (*f)(double), double x1, double x2);
double stochastic_grad_descent_1D(double (*f)(double), double x1, double x2);
#endif
