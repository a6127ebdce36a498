#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
E2E_CONFIGS="4:2048:1,4:512:0,4:256:0,4:1024:0,2:512:0,4:128:0,4:512:1,4:2048:1" timeout 300 $TR --nproc-per-node 8 --master-port 29833 tools/e2e_multi.py > gpurun_out/k_e2e_multi8.log 2> gpurun_out/k_e2e_multi8.err; echo "e2e_multi: $?"
cat gpurun_out/k_e2e_multi8.log
tail -n 3 gpurun_out/k_e2e_multi8.err
