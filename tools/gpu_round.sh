#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the ncu launch list of the same command and one full capture of
# the two passes. Run as: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
# tests written after round 1's GPU minutes were spent are opt-in (FVG_RUN_UNVERIFIED=1, tests/common.py) until they have
# passed once; they run first and on their own here, without -x, so that every one of them reports
( time FVG_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_post_r1_a_configs.py tests/test_post_r1_b_flow_conv.py tests/test_post_r1_c_reference_binding.py tests/test_post_r1_d_unsteady.py tests/test_post_r1_e_vortex.py -m gpu -q ) > $out/${tag}_pytest_gpu_new.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_gpu_new.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 300 python bench.py --numerics hllc-gg-bj --no-cpu-baseline > $out/${tag}_bench_n1_hllc_gg_bj.json 2> $out/${tag}_bench_n1_hllc_gg_bj.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cell_kernel|face_kernel' -s 6 -c 2 \
   -f -o $out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_full.log 2>&1
python tools/ncu_summary.py $out/${tag}_full.ncu-rep > $out/${tag}_ncu_summary.txt 2>&1
tail -3 $out/${tag}_pytest_gpu_new.log; tail -3 $out/${tag}_pytest_gpu.log; cat $out/${tag}_bench_n1.json; tail -2 $out/${tag}_bench_n1.err
# programmatic-dependent-launch build variant (never run in round 1): parity on a subset of the residual tests, then A/B timing
( timeout 900 make -C fvens_b200/csrc -j16 EXTRA=-DFVG_PDL OBJDIR=build_pdl TARGET=../variants_pdl.so > $out/${tag}_pdl_build.log 2>&1 \
  && FVENS_B200_LIB=$PWD/fvens_b200/variants_pdl.so timeout 600 python -m pytest tests/test_gpu_residual.py tests/test_gpu_solver.py -m gpu -x -q > $out/${tag}_pdl_pytest.log 2>&1; \
  tail -2 $out/${tag}_pdl_pytest.log; \
  timeout 600 bash tools/variant_sweep.sh "default:256 $PWD/fvens_b200/variants_pdl.so:256 default:256 $PWD/fvens_b200/variants_pdl.so:256" > $out/${tag}_pdl_sweep.txt 2>&1; cat $out/${tag}_pdl_sweep.txt )
