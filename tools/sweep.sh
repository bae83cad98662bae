#!/bin/bash
# kernel-variant sweep on the GPU box (each run is a fresh process: the switches are read once)
out=gpurun_out/sweep.log
: > $out
run() { env "$@" timeout 100 python tools/kbench.py --iters 150 --tag "$*" >> $out 2>&1; }
for pf in 0 2 4 8; do for v in 0 2; do run DIF_TC_P1_VARIANT=$v DIF_TC_P1_PREFETCH=$pf DIF_TC_P2_VARIANT=0 DIF_TC_P2_PREFETCH=0; done; done
for v in 1 2 4 6 7; do run DIF_TC_P1_VARIANT=2 DIF_TC_P1_PREFETCH=0 DIF_TC_P2_VARIANT=$v DIF_TC_P2_PREFETCH=0; done
for pf in 1 2; do for v in 0 6; do run DIF_TC_P1_VARIANT=2 DIF_TC_P1_PREFETCH=4 DIF_TC_P2_VARIANT=$v DIF_TC_P2_PREFETCH=$pf; done; done
cat $out
