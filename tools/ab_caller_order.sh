#!/bin/bash
# caller-ordered state: the one-tile-per-CTA gradient pass with the permutation gathers against the general template
out=gpurun_out; mkdir -p $out
run() { label=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --cell-order caller 2>$out/caller_order_err.log | tee $out/caller_order_$label.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', 'ms/step %.4f' % d['ms_per_step'], 'value %.3f' % d['value'], d.get('kernels_ms'), 'euler %.4f' % d['euler_step']['ms_per_step'], flush=True)
" || tail -3 $out/caller_order_err.log; }
{
run fastperm X=1
run general FVG_CELL_FASTPERM=0
timeout 300 python -m pytest tests/test_gpu_residual.py tests/test_gpu_cpp_surface.py -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee $out/perm12.log
