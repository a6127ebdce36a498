#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29831 tools/e2e_multi.py > gpurun_out/j_e2e_multi8.log 2> gpurun_out/j_e2e_multi8.err; echo "e2e_multi: $?" | tee gpurun_out/j.log
cat gpurun_out/j_e2e_multi8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29832 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/j_c3_weak8.json 2> gpurun_out/j_c3_weak8.err; echo "bench8: $?" | tee -a gpurun_out/j.log
python - <<'P'
import json
for f in ('j_c3_weak8',):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['config']['phases_ms_per_step'], d['config']['certification'], 'e2e %.4g %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), d['e2e']['phases_s'], 'gibbs %.4g'%d['gibbs']['value'])
P
tail -n 3 gpurun_out/j_e2e_multi8.err
