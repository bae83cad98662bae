#!/bin/bash
set -x
mkdir -p gpurun_out
cat VERSION_STAMP
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_pytest_g.log
grep -n "passed\|failed" gpurun_out/r2_pytest_g.log | tail -3; grep "^FAILED" gpurun_out/r2_pytest_g.log | head -8
bash tools/gpu_round2_prof.sh
