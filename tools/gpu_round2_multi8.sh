#!/bin/bash
# 8-GPU call: tests/test_gpu_multi.py (world 2, 4, 8), then bench.py at N = 8 (fused NVLink exchange + NCCL baseline), 4, 2, 1
set -x
mkdir -p gpurun_out
cat VERSION_STAMP
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -15 > gpurun_out/r2_multi_g8.log
tail -6 gpurun_out/r2_multi_g8.log
run() {  # N collective
  if [ "$1" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2_scale_n$1_$2.json 2> gpurun_out/r2_scale_n$1_$2.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $1 --steps 50 --warmup 10 --collective $2 > gpurun_out/r2_scale_n$1_$2.json 2> gpurun_out/r2_scale_n$1_$2.err
  fi
  python - <<P
import json
for l in open("gpurun_out/r2_scale_n$1_$2.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=$1 $2", "ms", round(d["ms_per_step"], 5), "G/s", round(d["value"] / 1e9, 3), "parity", d["parity"]["max"], d["parity"]["ok"],
              "| cfgB ms", round(d["cfg_b"]["ms_per_step"], 4), "frac", round(d["cfg_b"]["roofline_frac"], 3), "parity", d["cfg_b"]["parity"]["max"], "| e2e ms", round(d["e2e"]["ms_per_step"], 3))
P
  tail -2 gpurun_out/r2_scale_n$1_$2.err
}
run 8 nvlink
run 8 nccl
run 4 nvlink
run 2 nvlink
run 1 nvlink
