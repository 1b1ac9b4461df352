#!/bin/bash
# One GPU-box pass: parity tests, the bench lines, the ncu launch list of the bench command and one full capture of
# the two passes. Run as: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --durations=10 ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 300 python bench.py --numerics hllc-gg-bj --no-cpu-baseline > $out/${tag}_bench_n1_hllc_gg_bj.json 2> /dev/null
timeout 300 python bench.py --cell-order caller --no-cpu-baseline > $out/${tag}_bench_n1_caller_order.json 2> /dev/null
timeout 300 python bench.py --workload viscous --no-cpu-baseline > $out/${tag}_bench_n1_viscous.json 2> /dev/null
timeout 300 python bench.py --workload ogrid-weno --cells 6.25e6 --flux roe --no-cpu-baseline > $out/${tag}_bench_n1_ogrid_weno_roe.json 2> /dev/null
for fl in llf vanleer ausm hll hllc; do timeout 300 python bench.py --workload ogrid-weno --cells 6.25e6 --flux $fl --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_ogrid_weno_$fl.json 2> /dev/null; done
timeout 300 python bench.py --workload ogrid-weno --cells 6.25e6 --flux roe --weno-lambda 20 --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_ogrid_weno_roe_l20.json 2> /dev/null
for n in 512 2048 4096; do timeout 300 python bench.py --workload vortex --vortex-n $n --no-cpu-baseline --e2e-steps 1 --steps 30 > $out/${tag}_bench_n1_vortex_$n.json 2> /dev/null; done
for f in $out/${tag}_bench_n1*.json $out/${tag}_bench_ref.json; do tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('roofline') or {}).get('frac'), d.get('residual_roofline_frac'), (d.get('kernels_ms') or ''))" 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cell_kernel|face_kernel' -s 6 -c 2 \
   -f -o $out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_full.log 2>&1
python tools/ncu_summary.py $out/${tag}_full.ncu-rep > $out/${tag}_ncu_summary.txt 2>&1
head -44 $out/${tag}_ncu_summary.txt
