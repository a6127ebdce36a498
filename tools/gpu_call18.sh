#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c18.log
tail -n 3 gpurun_out/c18_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c18_smoke.log 2>&1; echo "smoke: $?" | tee -a gpurun_out/c18.log; tail -n 2 gpurun_out/c18_smoke.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c18_bench_c4.json 2> gpurun_out/c18_bench_c4.err; echo "bench c4: $?" | tee -a gpurun_out/c18.log
timeout 900 python bench.py > gpurun_out/c18_bench_c3.json 2> gpurun_out/c18_bench_c3.err; echo "bench c3 default: $?" | tee -a gpurun_out/c18.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 120 --csv --log-file gpurun_out/c18_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-iters 2 > gpurun_out/c18_ncu_launches.log 2>&1; echo "ncu launches: $?" | tee -a gpurun_out/c18.log
python - <<'P'
import json
for f in ('c18_bench_c4','c18_bench_c3'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], d['roofline'].get('frac'), d['roofline'].get('iteration_frac'), d.get('viterbi'), d['e2e'].get('value'), (d.get('gibbs') or {}).get('value'))
P
tail -n 3 gpurun_out/c18_bench_c3.err
