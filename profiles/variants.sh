python -m pytest tests/test_gpu_sweep.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -4
for lib in "" build/variants/libsmcb200_reg152.so build/variants/libsmcb200_reg144.so build/variants/libsmcb200_reg136.so; do
  SMCB200_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
done
