#!/usr/bin/env bash
# profiles/capture.sh <tag> [kernel-regex] [extra bench args...]
# Run on the GPU box (under gpurun): plain run first, then the per-launch time list, then ONE
# `--set full` capture of the named kernel (B200_PROFILING.md recipe).  Outputs in gpurun_out/.
set -uo pipefail
TAG="${1:-r01}"; KRE="${2:-k_sweep}"; shift 2 || true
CMD="python bench.py --steps 1 --warmup 3 --sweeps-per-step ${SPS:-40} --chains ${CHAINS:-1776} --no-e2e --no-cpu-baseline --no-extra $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP:-3} -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out | tail -8
