python -m pytest tests/test_gpu_allparticle.py -m gpu -q -x 2>&1 | tail -5
