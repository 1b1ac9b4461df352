out=gpurun_out; mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -q --durations=6 ) > $out/r02_pytest_gpu_final.log 2>&1
tail -4 $out/r02_pytest_gpu_final.log
timeout 300 python bench.py --workload viscous --no-cpu-baseline --steps 20 --warmup 5 > $out/r02_bench_n1_viscous.json 2>/dev/null
tail -1 $out/r02_bench_n1_viscous.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('viscous', d['ms_per_step'], d['value'], d['roofline']['frac'], d['residual_roofline_frac'], d['kernels_ms'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r02_bench_n1_final.json 2>/dev/null
tail -1 $out/r02_bench_n1_final.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
