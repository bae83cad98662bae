#!/bin/bash
# 1-GPU call: ncu launch list of bench.py + --set full captures of every shipped kernel family.  The reports are summarised ON the
# box (tools/ncu_summary.py -> markdown) and deleted: gpurun_out/ may only carry 64 MiB back.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg-b > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-200
prof() {   # name, kernel regex, launch-skip, launch-count
  DIF_PROFILE_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r2_prof_$1 python tools/profile_workloads.py > gpurun_out/r2_ncu_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$1.ncu-rep gpurun_out/r2_prof_$1.md "ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 python tools/profile_workloads.py" | tail -1
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
prof simple_fwd 'simple_fused|simple_lp' 3 3
prof simple_twopass 'reduce_tma|apply_tc' 0 12
prof sigmoid 'sigmoid_fwd_tc|sigmoid_bwd_tc|sigmoid_dq|sigmoid_dkv|sigmoid_fwd_kernel' 0 10
prof misc 'spmm_kernel|seg_fwd_warp|seg_bwd_warp|seg_bwd_fixup|reduce_kernel|apply_kernel' 0 12
ls -la gpurun_out | tail -12; du -sh gpurun_out
