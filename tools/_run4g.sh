N=4; tag=r02; out=gpurun_out; mkdir -p $out
run() { timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29512 --steps 50 --warmup 5 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
tail -1 $out/${tag}_bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N fused', d['ms_per_step'], d['value'], 'euler', d['euler_step']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'])"
run 29515 --workload vortex --vortex-n 4096 --steps 30 --warmup 5 --e2e-steps 1 > $out/${tag}_bench_n${N}_vortex_4096_strong.json 2>/dev/null
tail -1 $out/${tag}_bench_n${N}_vortex_4096_strong.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N vortex 4096 strong', d['ms_per_step'], d['value'], d['layout']['ghost_cells_on_rank0'])"
