#!/bin/bash
# Multi-GPU pass: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scale.sh N <tag> [check]'
# correctness on N real GPUs (tests/mgpu_check.py: bitwise against the single-GPU engine), then the bench line(s).
N=${1:-2}; tag=${2:-rXX}; check=${3:-check}; extras=${4:-extras}
out=gpurun_out; mkdir -p $out
if [ $check = check ]; then
( time timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py ) > $out/${tag}_mgpu_check_n$N.log 2>&1
grep -E "MGPU_CHECK" $out/${tag}_mgpu_check_n$N.log | tail -2
fi
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
tail -1 $out/${tag}_bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N fused', d['ms_per_step'], d['value'], 'euler', d['euler_step']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'])"
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 50 --warmup 5 --time-passes --e2e-steps 1 > $out/${tag}_bench_n${N}_passes.json 2>/dev/null
tail -1 $out/${tag}_bench_n${N}_passes.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N fused, passes timed', d['ms_per_step'], d.get('kernels_ms'))"
if [ $extras = extras ]; then
FVG_DIST=split timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --e2e-steps 1 > $out/${tag}_bench_n${N}_split.json 2> $out/${tag}_bench_n${N}_split.err
tail -1 $out/${tag}_bench_n${N}_split.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N split (round-1 schedule)', d['ms_per_step'], d['value'], 'euler', d['euler_step']['ms_per_step'])"
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 50 --warmup 5 --e2e-steps 1 --partition rcb 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N fused, rcb partition', d['ms_per_step'], 'euler', d['euler_step']['ms_per_step'], d['layout']['ghost_cells_on_rank0'])"
fi
