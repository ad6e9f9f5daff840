python -m pytest tests/test_gpu_equilibrium.py -m gpu -q -x -s > gpurun_out/eq.log 2>&1; grep -n "<E>" gpurun_out/eq.log | head
