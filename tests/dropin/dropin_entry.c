/* tests/dropin/dropin_entry.c — TEST INFRASTRUCTURE.
 * The drop-in (montecarlo-surfacer_b200/dropin/SMC.c, the reference's SMC.h API on libsmcb200) as a
 * shared library for ctypes, with the same test hooks as the compiled reference in oracle/_ref:
 * rand()/srand() are routed (by -Drand=replay_rand -Dsrand=replay_srand on the build line) to a
 * replayable integer stream, so the drop-in and the golden fixtures see the same random numbers. */
#include <stddef.h>
int replay_rand(void);
void replay_srand(unsigned s);
#include "SMC.c"

int ref_N(void) { return N; }
int ref_M(void) { return M; }
size_t ref_sizeof_sim(void) { return sizeof(struct Sim); }

/* sMC returns struct Sim by value; pointer trampoline for ctypes */
void ref_sMC(double L, double Lz, double T, double A, const double *W, const double *R0,
             int maxsteps, int gather_lapse, int eqsteps, struct Sim *out)
{
    *out = sMC(L, Lz, T, A, W, R0, maxsteps, gather_lapse, eqsteps);
}
