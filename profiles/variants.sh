python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batched', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['allparticle_kernel']['kernel_ms_per_step'])"
