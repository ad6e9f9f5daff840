#!/usr/bin/env bash
# profiles/capture_aux.sh <tag>: evidence for the kernels beside the headline one (run on the GPU box):
#  (1) the HBM-bound layout transposes, the evaluation and the gather at full size (8192 chains x N=256): duration and DRAM bytes
#      per launch (north_star: "achieved HBM GB/s for the move/accept kernels" - here move and accept are FUSED into the sweep
#      kernel, the only kernels that stream HBM are the AoS<->SoA transposes at the ABI);
#  (2) one --set full capture of the all-particle kernel at configs[4] (256 chains x N=4096).
set -uo pipefail
TAG="${1:-r02}"
mkdir -p gpurun_out
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
$CMD1 > /dev/null 2> gpurun_out/${TAG}_aux_plain.err || { echo "plain run failed"; tail -3 gpurun_out/${TAG}_aux_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_aos_to_soa|k_soa_to_aos|k_gather|k_evaluate_fast|k_chain_extent' --csv --log-file gpurun_out/${TAG}_aux_hbm.csv $CMD1 > gpurun_out/${TAG}_aux_ncu1.log 2>&1
CMD2="python bench.py --workload largeN --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-e2e"
$CMD2 > gpurun_out/${TAG}_largeN_plain.json 2> gpurun_out/${TAG}_largeN_plain.err || { echo "largeN plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_allparticle_fast -s 1 -c 1 -f -o gpurun_out/${TAG}_largeN_prof $CMD2 > gpurun_out/${TAG}_aux_ncu2.log 2>&1
ls -la gpurun_out | grep "${TAG}_" | tail -8
