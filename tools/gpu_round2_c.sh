#!/bin/bash
# 1-GPU call: 16-bit kernel first (short timeout), then launch-mode / L2-prefetch sweep of the one-kernel forward, timelines, ncu
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "16bit" 2>&1 | tail -15 > gpurun_out/r2_pytest_lp.log
tail -8 gpurun_out/r2_pytest_lp.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_simple_16bit_io 2>&1 | tail -8 > gpurun_out/r2_pytest_c.log
tail -4 gpurun_out/r2_pytest_c.log
python tools/kbench.py --membw | tee gpurun_out/r2_membw.log
( for cfg in "0 0" "0 3" "0 5" "0 7" "1 3" "2 0" "2 3" "2 5" "2 7"; do set -- $cfg
DIF_TC_LAUNCH=$1 DIF_TC_FUSED_PF_TILES=$2 timeout 200 python tools/kbench.py --iters 400 --only-fused --tag "launch=$1 pf=$2" 2>&1 | tail -1
done
for cfg in "0 0" "0 3" "2 3" "2 7"; do set -- $cfg
DIF_TC_LAUNCH=$1 DIF_TC_FUSED_PF_TILES=$2 timeout 200 python tools/kbench.py --iters 400 --dtype bf16 --tag "launch=$1 pf=$2" 2>&1 | tail -1
done
DIF_TC_LAUNCH=2 DIF_TC_FUSED_PF_TILES=3 timeout 200 python tools/kbench.py --iters 400 --dtype f16 --tag "launch=2 pf=3" 2>&1 | tail -1 ) | tee gpurun_out/r2_sweep_c.log
DIF_TC_DEBUG_TIMES=1 DIF_TC_FUSED_PF_TILES=3 timeout 200 python tools/kbench.py --iters 2 --only-fused > gpurun_out/r2_timeline_fused_c.log 2>&1
tail -14 gpurun_out/r2_timeline_fused_c.log
DIF_TC_DEBUG_TIMES=1 DIF_TC_FUSED_PF_TILES=3 timeout 200 python tools/kbench.py --iters 2 --dtype bf16 > gpurun_out/r2_timeline_lp_c.log 2>&1
tail -14 gpurun_out/r2_timeline_lp_c.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
tail -c 1200 gpurun_out/r2_bench_n1_c.json; tail -3 gpurun_out/r2_bench_n1_c.err
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'simple_fused|simple_lp' -s 6 -c 2 -o gpurun_out/r2_prof_fused python tools/kbench.py --iters 4 --only-fused > gpurun_out/r2_ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'simple_lp' -s 6 -c 2 -o gpurun_out/r2_prof_lp python tools/kbench.py --iters 4 --dtype bf16 > gpurun_out/r2_ncu_lp.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
