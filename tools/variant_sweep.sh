#!/bin/bash
# Times the residual with differently-built copies of the library (launch bounds / block sizes), tile sizes and
# environment knobs. usage: tools/variant_sweep.sh "<lib:tile[:ENV=val]> ..."   (lib = path of a .so or "default")
# Variants are built next to the default library (FVENS_B200_LIB selects one), e.g. other launch bounds:
#   make -C fvens_b200/csrc -j8 EXTRA="-DFVG_FACE_BLOCK=320" OBJDIR=build_320 TARGET=../variants_320.so
#   tools/variant_sweep.sh "default:256 fvens_b200/variants_320.so:256 default:256:FVG_PREFETCH_WAVES=0"
# This is how the A/B runs of profiles/r02_cell_kernel_ab.txt were made (round-1 library against the round-2 one on one box).
for spec in $1; do
  IFS=: read lib tile envs <<< "$spec"
  if [ "$lib" != "default" ]; then export FVENS_B200_LIB=$lib; else unset FVENS_B200_LIB; fi
  if [ -n "$envs" ]; then export $envs; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tile $tile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$spec', 'ms/step %.3f' % d['ms_per_step'], 'cell %.3f face %.3f' % (d['kernels_ms']['gradient_limiter_pass'], d['kernels_ms']['face_pass']), 'euler %.3f' % d['euler_step']['ms_per_step'], 'frac %.3f' % d['residual_roofline_frac'])
"
  if [ -n "$envs" ]; then unset ${envs%%=*}; fi
done
