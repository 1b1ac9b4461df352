#!/bin/bash
# A/B on one box: entry-capped tile cut (default) against the earlier cut (FVG_TILE_ENTRY_CAP=0) on the quad workloads and
# the headline; fixed-stride halo ids in the gradient pass (variants_hfix.so) against the default library.
run() { # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2>gpurun_out/ab_tile_cut_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d.get('kernels_ms',{})
print('$label', 'ms/step %.4f' % d['ms_per_step'], 'value %.3f' % d['value'], {a:round(b,4) for a,b in k.items()}, 'euler %.4f' % d['euler_step']['ms_per_step'], 'frac %.4f' % d['residual_roofline_frac'], flush=True)
" || tail -3 gpurun_out/ab_tile_cut_err.log
}
{
run headline_default X=1 --
run headline_hfix FVENS_B200_LIB=fvens_b200/variants_hfix.so --
run headline_default2 X=1 --
run ogrid_cap X=1 -- --workload ogrid-weno --cells 6.25e6 --flux roe
run ogrid_oldcut FVG_TILE_ENTRY_CAP=0 -- --workload ogrid-weno --cells 6.25e6 --flux roe
run vortex_cap X=1 -- --workload vortex --vortex-n 2500
run vortex_oldcut FVG_TILE_ENTRY_CAP=0 -- --workload vortex --vortex-n 2500
} 2>&1 | tee gpurun_out/ab_tile_cut.log
FVENS_B200_LIB=fvens_b200/variants_hfix.so timeout 600 python -m pytest tests/test_gpu_residual.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/ab_tile_cut.log
