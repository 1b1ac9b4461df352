mkdir -p gpurun_out
tag=r02h
for w in 0 1 2; do FVG_PREFETCH_WAVES=$w timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('prefetch waves $w', d['ms_per_step'], d['kernels_ms'])"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cell_kernel|face_kernel' -s 6 -c 2 \
   -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_ncu_summary.txt 2>&1
head -24 gpurun_out/${tag}_ncu_summary.txt
