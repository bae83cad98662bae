#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "sigmoid" 2>&1 | tail -15 > gpurun_out/r2_pytest_e_sigmoid.log
tail -6 gpurun_out/r2_pytest_e_sigmoid.log
timeout 1500 python -m pytest tests -m gpu -q -k "not sigmoid" 2>&1 | tail -30 > gpurun_out/r2_pytest_e.log
tail -12 gpurun_out/r2_pytest_e.log
rm -f gpurun_out/r2_stamps_e.csv
DIF_TC_DEBUG_TIMES=1 DIF_TC_DEBUG_CSV=gpurun_out/r2_stamps_e.csv timeout 200 python tools/kbench.py --iters 3 --only-fused > gpurun_out/r2_timeline_fused_e.log 2>&1
tail -16 gpurun_out/r2_timeline_fused_e.log
( for cfg in "0 0 -1" "2 0 -1" "3 0 -1" "2 2 -1" "2 4 -1" "2 0 32" "2 0 64" "2 0 96" "2 4 64"; do set -- $cfg
DIF_TC_LAUNCH=$1 DIF_TC_FUSED_PF_TILES=$2 DIF_TC_L2_PERSIST_MB=$3 timeout 200 python tools/kbench.py --iters 400 --only-fused --tag "launch=$1 pf=$2 persist=$3" 2>&1 | tail -2
done

DIF_TC_LAUNCH=2 timeout 200 python tools/kbench.py --iters 400 --only-fused --h 1 --tag "launch=2 H=1" 2>&1 | tail -1
DIF_TC_LAUNCH=2 timeout 200 python tools/kbench.py --iters 400 --only-fused --h 2 --tag "launch=2 H=2" 2>&1 | tail -1 ) | tee gpurun_out/r2_sweep_e.log
timeout 300 python bench.py --workload sigmoid_cora --steps 100 --warmup 5 | tee gpurun_out/r2_bench_sigmoid_cora.json | cut -c1-700
