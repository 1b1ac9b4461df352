#!/bin/bash
tag=r02h
out=gpurun_out
mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -q ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -5 $out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
for fl in roe llf vanleer ausm hll hllc; do timeout 200 python bench.py --workload ogrid-weno --cells 6.25e6 --flux $fl --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_ogrid_weno_$fl.json 2> /dev/null; done
timeout 200 python bench.py --workload ogrid-weno --cells 6.25e6 --flux roe --weno-lambda 20 --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_ogrid_weno_roe_l20.json 2> /dev/null
for n in 2048 4096; do timeout 200 python bench.py --workload vortex --vortex-n $n --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_vortex_$n.json 2> /dev/null; done
for f in $out/${tag}_bench_n1*.json; do tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('roofline') or {}).get('frac'), d.get('residual_roofline_frac'), (d.get('kernels_ms') or ''), (d.get('e2e') or {}).get('value'))" 2>/dev/null; done
