#!/bin/bash
# 1-GPU call near the end of the round: compute-sanitizer memcheck over the kernels added late (folded layer, batched graphs on tcgen05),
# --set full captures of seg_bwd_tc_kernel and the final layer_tc_kernel.
set -x
mkdir -p gpurun_out
cat VERSION_STAMP
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "v2_simple_tensor_core or projection_folded or layer_epilogue or shared_value_head" > gpurun_out/r2_sanitizer_memcheck_late_kernels.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_late_kernels.log
prof() {
  DIF_PROFILE_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r2_prof_$1 python tools/profile_workloads.py > gpurun_out/r2_ncu_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$1.ncu-rep gpurun_out/r2_prof_$1.md "ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 python tools/profile_workloads.py" | tail -1
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
prof segmented_tc 'seg_fwd_tc|seg_bwd_tc|seg_fixup_rows' 0 3
prof layer 'layer_tc_kernel|project_head|reduce_tma_kernel<1' 0 6
du -sh gpurun_out
