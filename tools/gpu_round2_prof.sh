#!/bin/bash
# 1-GPU call: ncu launch list of bench.py + one --set full capture covering every shipped kernel family
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg-b > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_ncu_bench.log | cut -c1-300
DIF_PROFILE_REPS=2 timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'simple_fused|simple_lp|reduce_tma|apply_tc|sigmoid_fwd_tc|sigmoid_bwd_tc|sigmoid_dq|sigmoid_dkv|sigmoid_fwd_kernel|spmm_kernel|seg_fwd_warp|seg_bwd_warp|seg_bwd_fixup|reduce_kernel|apply_kernel' \
  -o gpurun_out/r2_prof_all python tools/profile_workloads.py > gpurun_out/r2_ncu_all.log 2>&1
tail -3 gpurun_out/r2_ncu_all.log
ls -la gpurun_out/r2_prof_all.ncu-rep
