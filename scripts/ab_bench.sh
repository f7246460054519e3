for i in 1 2; do
for L in "" build_ab/libprev.so; do
  if [ -n "$L" ]; then export PN_B200_LIB=$PWD/$L; else unset PN_B200_LIB; fi
  timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_view']; print('lib=${L:-new}', round(d['value'],1), round(d['e2e']['value'],1), 'samp %.4f ref %.4f nerf %.4f gat %.4f'%(s['sampler_mlp'],s['refine_mlp'],s['nerf_mlp'],s['project_gather']), d['clocks']['sm_mhz'])"
done; done
