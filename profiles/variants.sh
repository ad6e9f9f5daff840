python -m pytest tests/test_gpu_allparticle.py tests/test_gpu_equilibrium.py -m gpu -q -x 2>&1 | tail -3
python bench.py --start droplet --kernel allparticle --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('droplet allparticle', d['value'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'], d['pairs_in_cutoff_frac'])"
python bench.py --kernel allparticle --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gas allparticle', d['value'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'])"
