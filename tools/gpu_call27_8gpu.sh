#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29821 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/i_c3_weak8_repair.json 2> gpurun_out/i_repair.err; echo "repair 1e-11: $?" | tee gpurun_out/i.log
BHMM_B200_REPAIR_TOL=0 timeout 600 $TR --nproc-per-node 8 --master-port 29822 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/i_c3_weak8_strict.json 2> gpurun_out/i_strict.err; echo "strict: $?" | tee -a gpurun_out/i.log
python - <<'P'
import json
for f in ('i_c3_weak8_repair','i_c3_weak8_strict'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['config']['phases_ms_per_step'], d['config']['certification'], 'e2e %.4g %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), 'gibbs %.4g'%d['gibbs']['value'])
P
tail -n 2 gpurun_out/i_repair.err
