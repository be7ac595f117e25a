#!/bin/bash
# GPU session of round 2: tests, bench, ncu captures of the HBM-bound kernels.  usage: tools/gpu_round2.sh <tag> [steps...]
TAG=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log ;;
    newtests) timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_guided.py -m gpu -q > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_new_$TAG.log; tail -15 gpurun_out/pytest_new_$TAG.log ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log ;;
    bench)   timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 800 gpurun_out/bench_ref_$TAG.json ;;
    ncu_ops) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"local_kernel|inpaint_kernel|fwht_rows|fwht_cols|wh_spec|sgemm_kernel|mul_table" -c 60 -o gpurun_out/prof_ops_$TAG -f python tools/profile_ops.py > gpurun_out/ncu_ops_$TAG.log 2>&1; tail -3 gpurun_out/ncu_ops_$TAG.log ;;
    ncu_gn)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_apply" -s 2 -c 8 -o gpurun_out/prof_gn_$TAG -f python tools/profile_ops.py sr4 > gpurun_out/ncu_gn_$TAG.log 2>&1; tail -3 gpurun_out/ncu_gn_$TAG.log ;;
    ncu_tc)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_" -s 60 -c 12 -o gpurun_out/prof_tc_$TAG -f python tools/profile_ops.py sr4 > gpurun_out/ncu_tc_$TAG.log 2>&1; tail -3 gpurun_out/ncu_tc_$TAG.log ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --profile-steps 2 > gpurun_out/ncu_bench_$TAG.log 2>&1; tail -2 gpurun_out/ncu_bench_$TAG.log ;;
    *) echo "running custom: $step"; timeout 900 bash -c "$step" ;;
  esac
done
