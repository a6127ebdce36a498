#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29841 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/m_c3_weak2.json 2> gpurun_out/m_weak2.err; echo "weak2: $?" | tee gpurun_out/m.log
timeout 600 $TR --nproc-per-node 2 --master-port 29842 bench.py --gpus 2 --scaling strong --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/m_c3_strong2.json 2> gpurun_out/m_strong2.err; echo "strong2: $?" | tee -a gpurun_out/m.log
python - <<'P'
import json
for f in ('m_c3_weak2','m_c3_strong2'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['config']['certification'], 'e2e %.4g %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), 'gibbs %.4g'%d['gibbs']['value'])
P
