#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29851 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/n_c3_weak4.json 2> gpurun_out/n_weak4.err; echo "weak4: $?" | tee gpurun_out/n.log
timeout 600 $TR --nproc-per-node 4 --master-port 29852 bench.py --gpus 4 --scaling strong --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/n_c3_strong4.json 2> gpurun_out/n_strong4.err; echo "strong4: $?" | tee -a gpurun_out/n.log
python - <<'P'
import json
for f in ('n_c3_weak4','n_c3_strong4'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['config']['certification'], 'e2e %.4g %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), 'gibbs %.4g'%d['gibbs']['value'])
P
