#!/bin/bash
# Multi-GPU pass on an N-GPU box: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scale.sh <tag> N'
tag=${1:-rXX}; n=${2:-2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > $out/${tag}_mgpu_check_n$n.log 2>&1; tail -3 $out/${tag}_mgpu_check_n$n.log
timeout 600 $TR --master-port 29512 bench.py --gpus $n --steps 100 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
FVG_FUSED_RECV=0 timeout 600 $TR --master-port 29513 bench.py --gpus $n --steps 100 --warmup 5 > $out/${tag}_bench_n${n}_recvkernel.json 2> $out/${tag}_bench_n${n}_recvkernel.err
MGPU_PARTITION=rcb timeout 300 $TR --master-port 29515 tests/mgpu_check.py > $out/${tag}_mgpu_check_rcb_n$n.log 2>&1; tail -2 $out/${tag}_mgpu_check_rcb_n$n.log
# coordinate-bisection partition (fewer ghost rows and neighbours; first timed in round 2)
timeout 600 $TR --master-port 29514 bench.py --gpus $n --steps 100 --warmup 5 --partition rcb > $out/${tag}_bench_n${n}_rcb.json 2> $out/${tag}_bench_n${n}_rcb.err
for f in $out/${tag}_bench_n$n.json $out/${tag}_bench_n${n}_recvkernel.json $out/${tag}_bench_n${n}_rcb.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', d['ms_per_step'], 'Gfaces/s', d['value'], 'euler ms', d['euler_step']['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'][:60])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tail -5 $out/${tag}_bench_n$n.err
