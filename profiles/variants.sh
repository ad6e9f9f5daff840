python bench.py --start droplet --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('droplet sweep', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['pairs_in_cutoff_frac'], d['acceptance'])"
python bench.py --start droplet --kernel allparticle --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('droplet allparticle', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['pairs_in_cutoff_frac'], d['acceptance'])"
