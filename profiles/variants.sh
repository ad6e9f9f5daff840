python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_ref.json 2>/dev/null
python bench.py --workload largeN --steps 3 --sweeps-per-step 10 --no-cpu-baseline > gpurun_out/final_largeN.json 2>/dev/null
python bench.py --workload grid --steps 2 --no-cpu-baseline > gpurun_out/final_grid.json 2>/dev/null
wc -c gpurun_out/final_*.json
