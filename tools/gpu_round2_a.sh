#!/bin/bash
# 2-GPU call: multi-GPU parity at HEAD, sanitizer on the cross-GPU exchange, bench at N=2 (fused NVLink vs NCCL)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15 > gpurun_out/r2_multi2.log
tail -5 gpurun_out/r2_multi2.log
export DIF_COMM_TIMEOUT_MS=120000
timeout 600 compute-sanitizer --target-processes all --tool memcheck --log-file gpurun_out/r2_san_multi_memcheck.%p.log \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sanitize_worker.py > gpurun_out/r2_san_multi_memcheck.out 2>&1
tail -3 gpurun_out/r2_san_multi_memcheck.out
timeout 600 compute-sanitizer --target-processes all --tool racecheck --log-file gpurun_out/r2_san_multi_racecheck.%p.log \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/sanitize_worker.py > gpurun_out/r2_san_multi_racecheck.out 2>&1
tail -3 gpurun_out/r2_san_multi_racecheck.out
timeout 600 compute-sanitizer --target-processes all --tool synccheck --log-file gpurun_out/r2_san_multi_synccheck.%p.log \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/sanitize_worker.py > gpurun_out/r2_san_multi_synccheck.out 2>&1
tail -3 gpurun_out/r2_san_multi_synccheck.out
unset DIF_COMM_TIMEOUT_MS
for coll in nvlink nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 2 --steps 100 --warmup 10 --collective $coll > gpurun_out/r2_bench_n2_$coll.json 2> gpurun_out/r2_bench_n2_$coll.err
tail -c 1500 gpurun_out/r2_bench_n2_$coll.json
tail -3 gpurun_out/r2_bench_n2_$coll.err
done
