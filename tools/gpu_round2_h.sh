#!/bin/bash
set -x
mkdir -p gpurun_out
cat VERSION_STAMP
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest_h.log
grep -n "passed\|failed" gpurun_out/r2_pytest_h.log | tail -3; grep "^FAILED" gpurun_out/r2_pytest_h.log | head -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
DIF_TC_DEBUG_TIMES=1 timeout 200 python tools/kbench.py --iters 3 --only-fused > gpurun_out/r2_timeline_fused.log 2>&1
DIF_TC_DEBUG_TIMES=1 timeout 200 python tools/kbench.py --iters 3 --dtype bf16 > gpurun_out/r2_timeline_lp.log 2>&1
tail -12 gpurun_out/r2_timeline_lp.log
( timeout 200 python tools/kbench.py --iters 400 --fused --tag "H=4" 2>&1 | tail -2
timeout 200 python tools/kbench.py --iters 400 --dtype bf16 --tag "H=4" 2>&1 | tail -1
for h in 1 2; do timeout 200 python tools/kbench.py --iters 400 --only-fused --h $h --tag "H=$h" 2>&1 | tail -1; timeout 200 python tools/kbench.py --iters 400 --dtype bf16 --h $h --tag "H=$h" 2>&1 | tail -1; done ) | tee gpurun_out/r2_kbench.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python - <<P
import json
for l in open("gpurun_out/r2_bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l); print("bench", d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("cold"), d["parity"]["max"], d["lp16"]["ms_per_step"], d["lp16"]["roofline"]["frac"], d["lp16"]["parity"], d["cfg_b"]["ms_per_step"], d["cfg_b"]["roofline_frac"], d["e2e"]["ms_per_step"], d["cpu_baseline"], d["torch_gpu_baseline"])
P
tail -2 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_n1_reference.json; cut -c1-300 gpurun_out/r2_bench_n1_reference.json
for w in sigmoid_cora layer segmented fwdbwd; do timeout 300 python bench.py --workload $w --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r2_bench_$w.json | cut -c1-500; done
