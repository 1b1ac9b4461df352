mkdir -p gpurun_out
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('plain', d['ms_per_step'], d['kernels_ms'], d['euler_step']['ms_per_step'], d['roofline']['frac'])"
(time timeout -k 10 900 python -m pytest tests/test_periodic.py tests/test_gpu_multirank.py tests/test_gpu_residual.py tests/test_gpu_solver.py -m gpu -q -x) > gpurun_out/r02g_pytest.log 2>&1; tail -30 gpurun_out/r02g_pytest.log
