#!/usr/bin/env bash
# profiles/multigpu.sh <ngpu> <tag>: the multi-GPU legs of the round on one box (run under `gpurun --gpus N`):
# weak and STRONG scaling of the headline batch, configs[3] (grid) and configs[4] (largeN) strong, and the C-side
# observable all-reduce test that needs two GPUs.  One JSON line per leg in gpurun_out/<tag>_*.json.
set -uo pipefail
N="${1:-2}"; TAG="${2:-r02}"
mkdir -p gpurun_out
run() { # name, args...
  local name="$1"; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" > "gpurun_out/${TAG}_${name}_${N}gpu.json" 2> "gpurun_out/${TAG}_${name}_${N}gpu.err" \
      || { echo "$name failed"; tail -5 "gpurun_out/${TAG}_${name}_${N}gpu.err"; }
  python - "$name" "gpurun_out/${TAG}_${name}_${N}gpu.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:14s} n_gpus {d['n_gpus']} scaling {d['scaling']:6s} value {d['value']:.4e} ms/step {d['ms_per_step']:.3f} kernel {d['kernel_ms_per_step']:.3f} "
          f"gather {d['gather_ms_per_step']:.3f} allreduce {d['allreduce_ms_per_step']:.3f} (exposed {d.get('allreduce_exposed_ms_per_step', 0):.3f}) e2e {d['e2e']['value']:.4e}")
except Exception as ex:
    print(sys.argv[1], "no line:", ex)
PY
}
python -m pytest tests/test_gpu_observables.py -q -m gpu -k "allreduce" 2>&1 | tail -3 | tee "gpurun_out/${TAG}_allreduce_test_${N}gpu.log"
run weak
run strong --scaling strong
run grid --workload grid
run largeN --workload largeN
