#!/bin/bash
# Round-2 profiler pass (one 1xB200 gpurun call): ncu launch list of the bench command, --set full of the three tcgen05 MLP
# kernels and of the HBM-side kernels of one step, and of refine_input_kernel on a 4032x3024 band (config 4).
# Numbers taken under ncu are never bench values; the summaries go to profiles/r02_*.
OUT=gpurun_out/r02_ncu
mkdir -p $OUT
B="python bench.py --no-extras --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $B --steps 2 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 9 -c 3 -o $OUT/prof_mlp $B --steps 1 --warmup 3 > $OUT/ncu_mlp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:refine_input_kernel|composite_scan_kernel|interval_refine_kernel" -s 9 -c 3 -o $OUT/prof_hbm $B --steps 1 --warmup 3 > $OUT/ncu_hbm.log 2>&1
timeout 600 python scripts/config4_band.py 378 3 > $OUT/config4_band.json 2> $OUT/config4_band.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:refine_input_kernel -s 1 -c 1 -o $OUT/prof_config4_gather python scripts/config4_band.py 378 2 > $OUT/ncu_config4.log 2>&1
for f in prof_mlp prof_hbm prof_config4_gather; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
done
ls -la $OUT; cat $OUT/config4_band.json
