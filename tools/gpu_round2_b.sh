#!/bin/bash
# 1-GPU call: full GPU test suite at HEAD, bench (one-kernel forward vs two launches), in-kernel timelines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest_b.log
tail -6 gpurun_out/r2_pytest_b.log
for path in fused twopass; do
timeout 600 python bench.py --steps 20 --warmup 5 --path $path > gpurun_out/r2_bench_n1_$path.json 2> gpurun_out/r2_bench_n1_$path.err
tail -c 600 gpurun_out/r2_bench_n1_$path.json; tail -3 gpurun_out/r2_bench_n1_$path.err
timeout 600 python bench.py --steps 200 --warmup 20 --path $path --no-cpu-baseline --no-cfg-b > gpurun_out/r2_bench_n1_${path}_200.json 2> gpurun_out/r2_bench_n1_${path}_200.err
python - <<P
import json
for l in open("gpurun_out/r2_bench_n1_${path}_200.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$path 200 steps", d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("cold"), d["parity"]["max"])
P
done
DIF_TC_DEBUG_TIMES=1 timeout 300 python tools/kbench.py --iters 3 --noref > gpurun_out/r2_timeline_twopass.log 2>&1
DIF_TC_DEBUG_TIMES=1 timeout 300 python tools/kbench.py --iters 3 --noref --fused > gpurun_out/r2_timeline_fused.log 2>&1
tail -40 gpurun_out/r2_timeline_fused.log
for rev in 0 1; do for hints in 0 1; do
DIF_TC_FUSED_REVERSE=$rev DIF_TC_P1_HINTS=$hints timeout 300 python tools/kbench.py --iters 300 --noref --fused --tag "rev=$rev hints=$hints" 2>&1 | tail -1
done; done | tee gpurun_out/r2_fused_sweep.log
