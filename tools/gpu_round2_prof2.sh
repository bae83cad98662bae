#!/bin/bash
# 1-GPU call: --set full captures of the kernels added late in round 2 (folded layer, batched graphs on tcgen05), summarised on the box.
set -x
mkdir -p gpurun_out
prof() {   # name, kernel regex, launch-skip, launch-count
  DIF_PROFILE_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r2_prof_$1 python tools/profile_workloads.py > gpurun_out/r2_ncu_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$1.ncu-rep gpurun_out/r2_prof_$1.md "ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 python tools/profile_workloads.py" | tail -1
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
prof layer 'layer_tc_kernel|project_head|project_finish' 0 10
prof segmented_tc 'seg_fwd_tc|seg_fwd_warp' 0 4
ls -la gpurun_out | tail -8; du -sh gpurun_out
