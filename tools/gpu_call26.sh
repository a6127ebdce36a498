#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/c26_sweep.log; : > $L
for v in "" libv_exprepl.so libv_atb.so libv_both.so ""; do
  if [ -n "$v" ]; then export BHMM_B200_LIB=$PWD/bhmm_b200/$v; else unset BHMM_B200_LIB; fi
  SWEEP_CHUNKS=2703 SWEEP_WARM=544 SWEEP_REPS=8 timeout 200 python tools/lane_sweep.py >> $L 2>&1
done
unset BHMM_B200_LIB
cat $L
T=gpurun_out/c26_transfer.log; : > $T
for kb in 1024 2048 4096; do BHMM_B200_STAGE_KB=$kb timeout 200 python tools/transfer_probe.py >> $T 2>&1; done
cat $T
