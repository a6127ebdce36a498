#!/bin/bash
# 8-GPU box: C3 weak + strong scaling, C5 (1e9 frames) time-sharded, C4 strong
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/m_c3_weak8.json 2> gpurun_out/m_c3_weak8.err; echo "c3 weak 8: $?" | tee gpurun_out/m.log
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 30 --warmup 5 --scaling strong > gpurun_out/m_c3_strong$n.json 2> gpurun_out/m_c3_strong$n.err; echo "c3 strong $n: $?" | tee -a gpurun_out/m.log
done
timeout 600 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --workload c5 --frames 1e9 --steps 3 --warmup 1 > gpurun_out/m_c5_8.json 2> gpurun_out/m_c5_8.err; echo "c5 8: $?" | tee -a gpurun_out/m.log
timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --workload c4 --scaling strong --steps 3 --warmup 1 > gpurun_out/m_c4_strong8.json 2> gpurun_out/m_c4_strong8.err; echo "c4 strong 8: $?" | tee -a gpurun_out/m.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/m_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['n_gpus'], d['scaling'], '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e', d['e2e'].get('value'), 'gibbs', (d.get('gibbs') or {}).get('value'), 'vit', (d.get('viterbi') or {}).get('frames_per_s'), d['config'].get('chunk'), d['config'].get('warm'))
P
tail -n 3 gpurun_out/m_*.err
