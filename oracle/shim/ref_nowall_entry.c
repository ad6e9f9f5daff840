/* oracle/shim/ref_nowall_entry.c — TEST INFRASTRUCTURE (oracle build only).
 *
 * The older bulk (3-D periodic, no wall) prototype, SMC_noMPI_noWall.c, as a
 * shared library.  Only its energy()/forces()/pressure() are used as parity
 * pins for BASELINE config 1 (its single-particle routines are inconsistent,
 * SURVEY.md §0-6).  build_ref.sh rewrites `#define N` in the scratch copy.
 */
#define main ref_nowall_main
#include "SMC_noMPI_noWall.c"
#undef main
int ref_N(void) { return N; }
