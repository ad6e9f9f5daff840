python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --no-cpu-baseline --no-e2e --thermalise 0 --kernel allparticle 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('allparticle half', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
SMCB_FULL_SHELL=1 python bench.py --steps 3 --no-cpu-baseline --no-e2e --thermalise 0 --kernel allparticle 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('allparticle full', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke()"
