N=2; tag=r02; out=gpurun_out; mkdir -p $out
run() { timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29515 --workload vortex --vortex-n 4096 --steps 30 --warmup 5 --e2e-steps 1 > $out/${tag}_bench_n${N}_vortex_4096_strong.json 2> $out/${tag}_bench_n${N}_vortex.err
tail -1 $out/${tag}_bench_n${N}_vortex_4096_strong.json | cut -c1-200; grep -v "^W\|^\*\|Setting OMP" $out/${tag}_bench_n${N}_vortex.err | tail -6
run 29516 --workload vortex --vortex-n 2896 --scaling weak --steps 30 --warmup 5 --e2e-steps 1 > $out/${tag}_bench_n${N}_vortex_2896_weak.json 2> $out/${tag}_bench_n${N}_vortex.err
tail -1 $out/${tag}_bench_n${N}_vortex_2896_weak.json | cut -c1-200
