#!/bin/bash
# Times the residual with differently-built copies of the library (launch bounds / block sizes), tile sizes and
# environment knobs. usage: tools/variant_sweep.sh "<lib:tile[:ENV=val]> ..."   (lib = path of a .so or "default")
# Variants are built next to the default library, e.g. the programmatic-dependent-launch build (pass B's prologue under
# pass A's tail and the next pass A's under pass B's; compiles, PREEXIT/ACQBULK in the SASS, never run yet):
#   make -C fvens_b200/csrc -j8 EXTRA=-DFVG_PDL OBJDIR=build_pdl TARGET=../variants_pdl.so
#   tools/variant_sweep.sh "default:256 fvens_b200/variants_pdl.so:256"
# and for N GPUs: FVENS_B200_LIB=$PWD/fvens_b200/variants_pdl.so torchrun ... bench.py --gpus N
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
