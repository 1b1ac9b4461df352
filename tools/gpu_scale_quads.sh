#!/bin/bash
# 2 GPUs after the tile-cut change: bitwise check against one GPU, configs[3] at 6.25 M cells per GPU, the headline mesh
out=gpurun_out; mkdir -p $out; tag=r02i
( time timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py ) > $out/${tag}_mgpu_check_n2.log 2>&1
grep -E "MGPU_CHECK" $out/${tag}_mgpu_check_n2.log | tail -2
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --workload ogrid-weno --cells 12.5e6 --flux roe --no-cpu-baseline --e2e-steps 1 > $out/${tag}_bench_n2_ogrid_weno_roe_12p5M.json 2> $out/${tag}_n2_ogrid.err
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_bench_n2.json 2> $out/${tag}_n2.err
for f in $out/${tag}_bench_n2*.json; do tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f'.split('/')[-1], d['ms_per_step'], d['value'], 'euler', d['euler_step']['ms_per_step'])"; done
