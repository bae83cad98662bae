#!/bin/bash
# 1-GPU call: full GPU suite (no -x), per-SM stamps, launch / prefetch / L2-persisting sweep of the one-kernel forward
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest_d.log
tail -12 gpurun_out/r2_pytest_d.log
rm -f gpurun_out/r2_stamps.csv
DIF_TC_DEBUG_TIMES=1 DIF_TC_DEBUG_CSV=gpurun_out/r2_stamps.csv timeout 200 python tools/kbench.py --iters 3 --only-fused > gpurun_out/r2_timeline_fused_d.log 2>&1
tail -16 gpurun_out/r2_timeline_fused_d.log
DIF_TC_DEBUG_TIMES=1 DIF_TC_DEBUG_CSV=gpurun_out/r2_stamps.csv timeout 200 python tools/kbench.py --iters 3 --dtype bf16 > gpurun_out/r2_timeline_lp_d.log 2>&1
tail -16 gpurun_out/r2_timeline_lp_d.log
( for cfg in "0 0 -1" "2 0 -1" "2 2 -1" "2 4 -1" "2 0 32" "2 0 64" "2 0 96" "2 4 96"; do set -- $cfg
DIF_TC_LAUNCH=$1 DIF_TC_FUSED_PF_TILES=$2 DIF_TC_L2_PERSIST_MB=$3 timeout 200 python tools/kbench.py --iters 400 --only-fused --tag "launch=$1 pf=$2 persist=$3" 2>&1 | tail -1
done
for cfg in "0 0 -1" "2 0 -1" "2 0 64"; do set -- $cfg
DIF_TC_LAUNCH=$1 DIF_TC_FUSED_PF_TILES=$2 DIF_TC_L2_PERSIST_MB=$3 timeout 200 python tools/kbench.py --iters 400 --dtype bf16 --tag "launch=$1 pf=$2 persist=$3" 2>&1 | tail -1
done
DIF_TC_LAUNCH=2 timeout 200 python tools/kbench.py --iters 400 --dtype f16 --tag "launch=2" 2>&1 | tail -1
DIF_TC_LAUNCH=2 timeout 200 python tools/kbench.py --iters 400 --only-fused --h 1 --tag "launch=2 H=1" 2>&1 | tail -1 ) | tee gpurun_out/r2_sweep_d.log
