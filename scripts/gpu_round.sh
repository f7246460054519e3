#!/bin/bash
# One 1xB200 box visit: parity tests, smoke, bench (both arms), parity report, reference-on-GPU comparison.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>      (ncu passes: scripts/ncu_r02.sh)
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
tail -5 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 400 $OUT/bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"
if [ -d baseline/_ref ]; then bash scripts/ref_gpu.sh run; cp gpurun_out/r02_ref_gpu.json gpurun_out/r02_parity.json $OUT/ 2>/dev/null; fi
ls -la $OUT
