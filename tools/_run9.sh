mkdir -p gpurun_out
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('default', d['ms_per_step'], d['kernels_ms'], d['euler_step']['ms_per_step'])"
(time timeout -k 10 1300 python -m pytest tests/test_gpu_residual.py tests/test_gpu_multirank.py tests/test_gpu_scale.py tests/test_periodic.py -m gpu -q -x) > gpurun_out/r02j_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02j_pytest_gpu.log
