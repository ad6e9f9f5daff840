python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for cl in 1 4; do
SMCB_CLUSTER=$cl python bench.py --workload largeN --steps 3 --sweeps-per-step 10 --no-cpu-baseline --no-e2e --chains 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cluster $cl', d['value'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'])"
done
python bench.py --steps 3 --no-cpu-baseline --no-e2e --thermalise 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batched', d['value'], d['kernel_ms_per_step'], d['gather_ms_per_step'], d['roofline']['frac'])"
