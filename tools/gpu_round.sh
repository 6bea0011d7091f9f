#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, smoke, bench lines (cfg2, cfg4, reference arm), the ncu launch
# list of whole steps, one `ncu --set full` capture of the dominant kernel, and the standalone harness timeline.
# Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py condenses the .ncu-rep for profiles/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 1500 $O/bench_cfg2.json
DMH_TUNING=tile=0 timeout 600 python bench.py --no-cpu-baseline > $O/bench_cfg2_scalar.json 2>/dev/null; tail -c 600 $O/bench_cfg2_scalar.json
timeout 600 python bench.py --workload cfg4 --steps 10 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 1500 $O/bench_cfg4.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2>&1; tail -c 800 $O/bench_ref.json
if [ -x tools/tile_bench ]; then
  tools/tile_bench 64 1 320 576 32 20 > $O/tile_bench.txt 2>&1
  tools/tile_bench 64 1 320 576 32 20 1 >> $O/tile_bench.txt 2>&1
  tools/tile_bench 128 3 512 512 32 10 >> $O/tile_bench.txt 2>&1
  tools/tile_bench 128 3 512 512 32 10 tile=2 >> $O/tile_bench.txt 2>&1
  tools/tile_bench 16 3 1080 1920 64 10 1 >> $O/tile_bench.txt 2>&1
  cat $O/tile_bench.txt
fi
[ -x tools/tile_bench_dbg ] && tools/tile_bench_dbg 64 1 320 576 32 5 > $O/tile_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_step.csv \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_tile_kernel -s 6 -c 1 -f -o $O/tile_full \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
