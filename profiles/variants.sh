python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --no-cpu-baseline --thermalise 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batched', d['value'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
python bench.py --workload grid --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/grid.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('grid', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['workload'])"; tail -2 gpurun_out/grid.err
