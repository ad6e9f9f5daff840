python -m pytest tests/test_gpu_allparticle.py tests/test_gpu_equilibrium.py tests/test_gpu_observables.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --no-cpu-baseline --no-e2e --thermalise 0 --kernel allparticle 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('allparticle', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
