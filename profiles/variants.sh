for lib in "" build/variants/libsmcb200_reg152.so build/variants/libsmcb200_reg144.so; do
  SMCB200_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 3 --thermalise 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])"
done
