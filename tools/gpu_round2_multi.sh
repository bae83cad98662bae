#!/bin/bash
# multi-GPU call (run with gpurun --gpus G): tests/test_gpu_multi.py for every world size the box offers, then bench.py at N = G
# (weak scaling of config A with per-rank fp64 parity + config B strong scaling), fused NVLink exchange and the NCCL baseline
set -x
G=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -15 > gpurun_out/r2_multi_g$G.log
tail -6 gpurun_out/r2_multi_g$G.log
for coll in nvlink nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $G --steps 50 --warmup 10 --collective $coll > gpurun_out/r2_bench_n${G}_$coll.json 2> gpurun_out/r2_bench_n${G}_$coll.err
python - <<P
import json
for l in open("gpurun_out/r2_bench_n${G}_$coll.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=$G $coll", "ms", d["ms_per_step"], "value G/s", d["value"] / 1e9, "parity", d["parity"]["max"], d["parity"]["ok"],
              "cfgB ms", d["cfg_b"]["ms_per_step"], "cfgB frac", d["cfg_b"]["roofline_frac"], "cfgB parity", d["cfg_b"]["parity"]["max"], "e2e ms", d["e2e"]["ms_per_step"])
P
tail -2 gpurun_out/r2_bench_n${G}_$coll.err
done
