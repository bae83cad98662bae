#!/bin/bash
set -x
mkdir -p gpurun_out
cat VERSION_STAMP 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest_f.log
grep -n "passed\|failed" gpurun_out/r2_pytest_f.log | tail -3; grep "^FAILED" gpurun_out/r2_pytest_f.log | head -5
rm -f gpurun_out/r2_stamps_f.csv
DIF_TC_DEBUG_TIMES=1 DIF_TC_DEBUG_CSV=gpurun_out/r2_stamps_f.csv timeout 200 python tools/kbench.py --iters 3 --only-fused > gpurun_out/r2_timeline_fused_f.log 2>&1
tail -12 gpurun_out/r2_timeline_fused_f.log
DIF_TC_DEBUG_TIMES=1 timeout 200 python tools/kbench.py --iters 3 --dtype bf16 > gpurun_out/r2_timeline_lp_f.log 2>&1
tail -12 gpurun_out/r2_timeline_lp_f.log
( for cfg in "0" "2"; do
DIF_TC_LAUNCH=$cfg timeout 200 python tools/kbench.py --iters 400 --only-fused --tag "launch=$cfg" 2>&1 | tail -1
DIF_TC_LAUNCH=$cfg timeout 200 python tools/kbench.py --iters 400 --dtype bf16 --tag "launch=$cfg" 2>&1 | tail -1
done
timeout 200 python tools/kbench.py --iters 400 --fused --tag "two-pass vs fused" 2>&1 | tail -2 ) | tee gpurun_out/r2_sweep_f.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err
python - <<P
import json
for l in open("gpurun_out/r2_bench_n1_f.json"):
    if l.startswith("{"):
        d = json.loads(l); print("bench", d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("cold"), d["parity"]["max"], d["lp16"]["ms_per_step"], d["lp16"]["roofline"]["frac"], d["cfg_b"]["ms_per_step"], d["cfg_b"]["roofline_frac"], d["e2e"]["ms_per_step"])
P
tail -2 gpurun_out/r2_bench_n1_f.err
for w in sigmoid_cora layer segmented fwdbwd; do timeout 300 python bench.py --workload $w --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r2_bench_$w.json | cut -c1-400; done
bash tools/gpu_round2_prof.sh
