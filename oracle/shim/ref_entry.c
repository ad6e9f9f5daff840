/* oracle/shim/ref_entry.c — TEST INFRASTRUCTURE (oracle build only).
 *
 * Single translation unit of the UNMODIFIED reference: main.c textually
 * includes SMC.c -> SMC.h -> matematicose.c, misccose.c (main.c:2, SMC.c:15,
 * SMC.h:19-20).  Its main() is renamed so the result can be a shared library.
 * build_ref.sh compiles this from a scratch copy of /root/reference under /tmp
 * in which only the `#define N` / `#define M` lines of SMC.h are rewritten
 * (sizes are compile-time macros in the reference, SMC.h:26-29).
 */
#define main ref_main
#include "main.c"
#undef main

/* sizes this build was specialised for, so a test can assert it loaded the right one */
int ref_N(void) { return N; }
int ref_M(void) { return M; }
double ref_a0(void) { return a0; }
double ref_b0(void) { return b0; }
double ref_cutoff(void) { double L = 0.0; (void)L; return LJ_CUTOFF; }
size_t ref_sizeof_sim(void) { return sizeof(struct Sim); }

/* sMC returns `struct Sim` by value (SMC.h:92); ctypes-friendly trampoline */
void ref_sMC(double L, double Lz, double T, double A, const double *W, const double *R0,
             int maxsteps, int gather_lapse, int eqsteps, struct Sim *out)
{
    *out = sMC(L, Lz, T, A, W, R0, maxsteps, gather_lapse, eqsteps);
}
