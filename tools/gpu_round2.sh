#!/bin/bash
# GPU session of round 2: tests, bench, ncu captures of the HBM-bound kernels.  usage: tools/gpu_round2.sh <tag> [steps...]
TAG=$1; shift
mkdir -p gpurun_out
# .ncu-rep files are too large to bring back (64 MiB cap on gpurun_out): export the raw metric page (and, for the convolution
# kernel, the per-instruction source page) as CSV on the box and drop the report
ncu_export() {   # <report stem> [source]
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  if [ "$2" = "source" ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null; gzip -f gpurun_out/$1.source.csv; fi
  rm -f gpurun_out/$1.ncu-rep
}
for step in "$@"; do
  case $step in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log ;;
    newtests) timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_guided.py -m gpu -q > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_new_$TAG.log; tail -15 gpurun_out/pytest_new_$TAG.log ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log ;;
    bench)   timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 800 gpurun_out/bench_ref_$TAG.json ;;
    ncu_ops) for op in sr4 inpaint wh deblur; do timeout 600 ncu --set full --clock-control none -k regex:"local_kernel|inpaint_kernel|fwht_rows|fwht_cols|wh_spec|sgemm_kernel|mul_table|final_" -c 10 -o gpurun_out/prof_op_${op}_$TAG -f python tools/profile_ops.py $op > gpurun_out/ncu_op_${op}_$TAG.log 2>&1; tail -1 gpurun_out/ncu_op_${op}_$TAG.log; ncu_export prof_op_${op}_$TAG; done ;;
    ncu_gn)  DDNM_GN_FUSED=0 timeout 600 ncu --set full --clock-control none -k regex:"gn_apply" -s 2 -c 5 -o gpurun_out/prof_gn_$TAG -f python tools/profile_ops.py sr4 > gpurun_out/ncu_gn_$TAG.log 2>&1; tail -1 gpurun_out/ncu_gn_$TAG.log; ncu_export prof_gn_$TAG ;;
    ncu_tc)  DDNM_GN_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gn" -s 16 -c 3 -o gpurun_out/prof_tcgn_$TAG -f python tools/profile_ops.py sr4 > gpurun_out/ncu_tc_$TAG.log 2>&1; tail -1 gpurun_out/ncu_tc_$TAG.log; ncu_export prof_tcgn_$TAG source ;;
    timing)  timeout 600 python tests/diag/gn_conv_diag.py timing 0 > gpurun_out/gn_timing_$TAG.log 2>&1; cat gpurun_out/gn_timing_$TAG.log | tail -28 ;;
    ab)      DDNM_GN_FUSED=0 tools/gpu_session.sh ab_unfused_$TAG unet_bench:celeba:16:5 | tail -25; DDNM_GN_FUSED=1 tools/gpu_session.sh ab_fused_$TAG unet_bench:celeba:16:5 openai_bench:8:3 | tail -50 ;;
    abpdl)   DDNM_PDL=0 tools/gpu_session.sh ab_nopdl_$TAG unet_bench:celeba:16:5 | grep "unet bench"; DDNM_PDL=1 tools/gpu_session.sh ab_pdl_$TAG unet_bench:celeba:16:5 openai_bench:8:3 | grep "bench" ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --profile-steps 2 > gpurun_out/ncu_bench_$TAG.log 2>&1; tail -2 gpurun_out/ncu_bench_$TAG.log ;;
    ncu_head) timeout 600 ncu --set full --clock-control none -k regex:"head_conv|conv_small_cin" -s 2 -c 2 -o gpurun_out/prof_head_$TAG -f python tools/ncu_fwd.py > gpurun_out/ncu_head_$TAG.log 2>&1; tail -1 gpurun_out/ncu_head_$TAG.log; ncu_export prof_head_$TAG ;;
    ncu_fwd:*) # ncu_fwd:<index among the 112 tensor-core launches of a forward>[:source]
             IFS=: read -ra pp <<< "$step"; idx=${pp[1]}
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s $((112 + idx)) -c 1 -o gpurun_out/prof_fwd${idx}_$TAG -f python tools/ncu_fwd.py > gpurun_out/ncu_fwd${idx}_$TAG.log 2>&1; tail -1 gpurun_out/ncu_fwd${idx}_$TAG.log; ncu_export prof_fwd${idx}_$TAG source ;;
    *) echo "running custom: $step"; timeout 900 bash -c "$step" ;;
  esac
done
