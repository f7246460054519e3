#!/bin/bash
# A/B of two builds of the library on the SAME box (box-to-box spread is +-1.5 %, more than most kernel changes):
# alternates bench.py between the in-tree build and build_ab/libprev.so (PN_B200_LIB override, see _abi.py) and prints
# kernel-only / end-to-end Mrays/s and the MLP + gather stage times per view.
#
#   # build the "previous" library from a commit's sources (nvcc cross-compiles here, the .so travels with gpurun):
#   mkdir -p /tmp/prev/x/y/csrc /tmp/prev/x/include build_ab
#   for f in api.cu elementwise.cu gather.cu mlp_f32.cu mlp_prog.cu mlp_tc.cu common.cuh tc.cuh; do git show <rev>:pronerf_b200/csrc/$f > /tmp/prev/x/y/csrc/$f; done
#   git show <rev>:include/pronerf_b200.h > /tmp/prev/x/include/pronerf_b200.h
#   (cd /tmp/prev/x/y/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
#        -o $OLDPWD/build_ab/libprev.so api.cu elementwise.cu gather.cu mlp_f32.cu mlp_prog.cu mlp_tc.cu)
#   gpurun --timeout 400 -- 'bash scripts/ab_bench.sh'
# (the loader binds every symbol of include/pronerf_b200.h: <rev> must not be older than the newest entry point, pn_debug_tc_clock;
#  scripts/tc_power.py does the same A/B as trains of launches at the board's power limit, scripts/tc_timeline.py in cycles)
for i in 1 2; do
for L in "" build_ab/libprev.so; do
  if [ -n "$L" ]; then export PN_B200_LIB=$PWD/$L; else unset PN_B200_LIB; fi
  timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step_rank0']; print('lib=${L:-new}', round(d['value'],1), round(d['e2e']['value'],1), 'samp %.4f ref %.4f nerf %.4f gat %.4f'%(s['sampler_mlp'],s['refine_mlp'],s['nerf_mlp'],s['project_gather']), d['clocks']['sm_mhz'])"
done; done
