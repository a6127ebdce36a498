#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --workload c5 --frames 1e9 --steps 3 --warmup 1 > gpurun_out/f_c5_8.json 2> gpurun_out/f_c5_8.err; echo "c5 8: $?" | tee gpurun_out/f2.log
python - <<'P'
import json
for l in open('gpurun_out/f_c5_8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['viterbi'], d['config']['certification'], d['config']['worst_border_mismatch'])
P
tail -n 4 gpurun_out/f_c5_8.err
