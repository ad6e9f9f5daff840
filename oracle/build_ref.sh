#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.
#
# Compiles the real reference (read-only under /root/reference) into shared
# libraries under oracle/_ref/ so tests can call its own routines as the parity
# oracle.  Nothing is copied into the repository: the sources are copied to a
# scratch directory under /tmp, the compile-time size macros are rewritten there
# (SMC.h:26-29, SMC_noMPI_noWall.c:16-18), and only the .so files come back.
# oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.
#
# Flags: -std=gnu11 -O2 -ffp-contract=off and NO -march=x86-64-v3:
# FMA contraction alone changes trajectories (the reference is chaotic at 1e-16).
# An additional -O3 -march=x86-64-v3 build (AVX2+FMA; not "native", the .so travels to another host) (suffix _fast) is the courtesy CPU
# baseline number, never a parity oracle.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SMCB_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/SMC.c" ]; then
    echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
SCRATCH="$(mktemp -d /tmp/smcb_ref_build.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT
CFLAGS_PARITY="-std=gnu11 -O2 -ffp-contract=off -fPIC -shared -w"
CFLAGS_FAST="-std=gnu11 -O3 -march=x86-64-v3 -fPIC -shared -w"

cc -std=gnu11 -O2 -fPIC -c "$HERE/shim/oracle_rand.c" -o "$SCRATCH/oracle_rand.o"

build_wall() {  # N M
    local n="$1" m="$2" d="$SCRATCH/w_${1}_${2}"
    mkdir -p "$d"
    cp "$REF"/SMC.c "$REF"/SMC.h "$REF"/main.c "$REF"/matematicose.c "$REF"/matematicose.h "$d/"
    cp "$HERE/shim/misccose.c" "$HERE/shim/fftw3.h" "$HERE/shim/ref_entry.c" "$d/"
    sed -i -e "s/^#define N 108\s*$/#define N ${n}/" -e "s/^#define M 3\s*$/#define M ${m}/" "$d/SMC.h"
    grep -q "^#define N ${n}\$" "$d/SMC.h" && grep -q "^#define M ${m}\$" "$d/SMC.h"
    local tag="N${n}_M${m}"
    gcc $CFLAGS_PARITY -I"$d" -Drand=oracle_rand -Dsrand=oracle_srand \
        "$d/ref_entry.c" "$SCRATCH/oracle_rand.o" -lm -o "$OUT/libref_${tag}.so"
    gcc $CFLAGS_FAST -I"$d" -Drand=oracle_rand -Dsrand=oracle_srand \
        "$d/ref_entry.c" "$SCRATCH/oracle_rand.o" -lm -o "$OUT/libref_${tag}_fast.so"
}

build_nowall() {  # N
    local n="$1" d="$SCRATCH/nw_${1}"
    mkdir -p "$d"
    cp "$REF"/SMC_noMPI_noWall.c "$d/"
    cp "$HERE/shim/fftw3.h" "$HERE/shim/ref_nowall_entry.c" "$d/"
    sed -i -e "s/^#define N 32\s*$/#define N ${n}/" "$d/SMC_noMPI_noWall.c"
    grep -q "^#define N ${n}\$" "$d/SMC_noMPI_noWall.c"
    gcc $CFLAGS_PARITY -I"$d" -Drand=oracle_rand -Dsrand=oracle_srand \
        "$d/ref_nowall_entry.c" "$SCRATCH/oracle_rand.o" -lm -o "$OUT/libref_nowall_N${n}.so"
}

for cfg in "32 3" "108 3" "256 3" "500 3" "108 4" "4096 3"; do
    set -- $cfg
    build_wall "$1" "$2"
done
build_nowall 108
build_nowall 32
ls -la "$OUT"
