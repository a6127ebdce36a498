#!/bin/bash
mkdir -p gpurun_out
for w in c5 n32 n3 c1 c2; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/c13_bench_$w.json 2> gpurun_out/c13_bench_$w.err; echo "bench $w: $?" | tee -a gpurun_out/c13.log
done
python - <<'P'
import json
for w in ('c5','n32','n3','c1','c2'):
    for l in open('gpurun_out/c13_bench_%s.json'%w):
        if l.startswith('{'):
            d=json.loads(l); print(w, '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('frac'), d['roofline'].get('iteration_frac'), d['config'].get('certification'), (d.get('e2e') or {}).get('value'), (d.get('gibbs') or {}).get('value'), (d.get('viterbi') or {}).get('frames_per_s'), (d.get('cpu_baseline') or {}).get('value'))
P
tail -n 4 gpurun_out/c13_bench_*.err
