mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
tools/ubench_ffma2 > $O/r2_ubench_ffma2.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --steps 30 > $O/r2_bench0.json 2> $O/r2_bench0.err; tail -c 1200 $O/r2_bench0.json
cat $O/r2_ubench_ffma2.txt
