#!/bin/bash
# 8-GPU box, final: C5 1e9 frames (forward-backward + Viterbi across shards), C3 weak/strong with the final code, gloo-free NCCL estimator test
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --workload c5 --frames 1e9 --steps 3 --warmup 1 > gpurun_out/f_c5_8.json 2> gpurun_out/f_c5_8.err; echo "c5 8: $?" | tee gpurun_out/f.log
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/f_c3_weak8.json 2> gpurun_out/f_c3_weak8.err; echo "c3 weak 8: $?" | tee -a gpurun_out/f.log
timeout 600 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 30 --warmup 5 --scaling strong > gpurun_out/f_c3_strong8.json 2> gpurun_out/f_c3_strong8.err; echo "c3 strong 8: $?" | tee -a gpurun_out/f.log
timeout 600 $TR --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --workload c4 --scaling strong --steps 3 --warmup 1 > gpurun_out/f_c4_strong8.json 2> gpurun_out/f_c4_strong8.err; echo "c4 strong 8: $?" | tee -a gpurun_out/f.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['n_gpus'], d['scaling'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e', d['e2e'].get('value'), 'gibbs', (d.get('gibbs') or {}).get('value'), 'vit', d.get('viterbi'))
P
tail -n 4 gpurun_out/f_c5_8.err
