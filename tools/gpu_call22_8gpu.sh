#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/g_c3_weak8.json 2> gpurun_out/g_c3_weak8.err; echo "c3 weak 8: $?" | tee gpurun_out/g.log
timeout 300 $TR --nproc-per-node 8 --master-port 29712 tools/allreduce_probe.py > gpurun_out/g_allreduce_probe.log 2>&1; echo "allreduce probe: $?" | tee -a gpurun_out/g.log
python - <<'P'
import json
for l in open('gpurun_out/g_c3_weak8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['phases_ms_per_step'], d['roofline']['all_kernels_ms'], d['config']['certification'])
P
tail -n 6 gpurun_out/g_allreduce_probe.log
