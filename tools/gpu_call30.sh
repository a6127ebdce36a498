#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/c30_transfer.log; : > $T
timeout 300 python -m pytest tests/test_transfer_cuda.py -x -q 2>&1 | tail -n 2 | tee -a $T
for nt in 1 0 1 0; do echo "NT_COPY=$nt" >> $T; BHMM_B200_NT_COPY=$nt PROBE_THREADS=2,4,8 timeout 200 python tools/transfer_probe.py >> $T 2>&1; done
cat $T
