#!/bin/bash
# Builder-side: run the UNMODIFIED reference on the same B200 as this package and compare (scripts/ref_gpu.py).
#   scripts/ref_gpu.sh stage   -- here (where /root/reference exists): copy the three hot-path files into the git-ignored
#                                 baseline/_ref/ so that they travel with the gpurun snapshot (never committed)
#   scripts/ref_gpu.sh run     -- on the GPU box (inside a gpurun command): reference half, then this package's half
set -e
cd "$(dirname "$0")/.."
case "$1" in
  stage)
    mkdir -p baseline/_ref
    for f in run_S_eS_eN_alter_trt.py run_nerf_helpers.py inverse_warp.py run_S_eS_eN_alter_base.py; do
      cp /root/reference/$f baseline/_ref/$f
    done
    ls -la baseline/_ref ;;
  run)
    out=/tmp/pn_ref_gpu
    mkdir -p $out gpurun_out
    python scripts/ref_gpu.py ref $out  > gpurun_out/r02_ref_gpu_ref.log 2>&1 || { tail -30 gpurun_out/r02_ref_gpu_ref.log; exit 1; }
    python scripts/ref_gpu.py ours $out > gpurun_out/r02_ref_gpu_ours.log 2>&1 || { tail -30 gpurun_out/r02_ref_gpu_ours.log; exit 1; }
    python tests/parity_report.py gpurun_out/r02_parity.json $out > gpurun_out/r02_parity.log 2>&1 || { tail -30 gpurun_out/r02_parity.log; exit 1; }
    tail -3 gpurun_out/r02_ref_gpu_ours.log ;;
  *) echo "usage: $0 stage|run"; exit 2 ;;
esac
