#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29811 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/h_c3_weak8_opt.json 2> gpurun_out/h_opt.err; echo "opt: $?" | tee gpurun_out/h.log
BHMM_B200_OPTIMISTIC=0 timeout 600 $TR --nproc-per-node 8 --master-port 29812 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/h_c3_weak8_pess.json 2> gpurun_out/h_pess.err; echo "pess: $?" | tee -a gpurun_out/h.log
python - <<'P'
import json
for f in ('h_c3_weak8_opt','h_c3_weak8_pess'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['config']['phases_ms_per_step'], d['e2e']['value'])
P
