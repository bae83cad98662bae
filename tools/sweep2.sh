#!/bin/bash
out=gpurun_out/sweep2.log
: > $out
run() { env "$@" timeout 100 python tools/kbench.py --iters 150 --tag "$*" >> $out 2>&1; }
for p1 in 2 6; do for pf in 0 4; do for p2 in 0 6; do for pf2 in 0 1; do run DIF_TC_P1_VARIANT=$p1 DIF_TC_P1_PREFETCH=$pf DIF_TC_P2_VARIANT=$p2 DIF_TC_P2_PREFETCH=$pf2; done; done; done; done
cat $out
