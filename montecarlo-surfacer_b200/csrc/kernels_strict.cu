// kernels_strict.cu — STRICT instantiations.  MUST be compiled with --fmad=false:
// the code below mirrors the reference's C expressions and relies on the compiler
// never contracting a*b+c into an FMA (the reference oracle is built with
// -ffp-contract=off).  The Makefile enforces the flag; the static_assert-style
// guard below catches a wrong build line.
#ifndef SMCB_FMAD_OFF
#error "kernels_strict.cu must be built with --fmad=false -DSMCB_FMAD_OFF"
#endif
#include "launch.h"
#define SMCB_TU_IS_STRICT 1
#define SMCB_TU_STRICT true
#define SMCB_TU_SUFFIX strict
#include "launchers.inl"
