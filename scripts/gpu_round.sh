#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, one full capture of the dominant kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
tail -6 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 6 -c 3 -o $OUT/prof_mlp \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:refine_input_kernel|composite_scan_kernel|interval_refine_kernel" -s 9 -c 3 -o $OUT/prof_gather \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_gather.log 2>&1
fi
ls -la $OUT
