#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/f.log
tail -n 4 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke: $?" | tee -a gpurun_out/f.log; tail -n 1 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench_c3.json 2> gpurun_out/f_bench_c3.err; echo "bench: $?" | tee -a gpurun_out/f.log
timeout 900 python bench.py --impl reference > gpurun_out/f_bench_c3_reference.json 2> gpurun_out/f_bench_c3_reference.err; echo "reference arm: $?" | tee -a gpurun_out/f.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 160 --csv --log-file gpurun_out/f_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-iters 2 > gpurun_out/f_ncu_launches.log 2>&1; echo "ncu launches: $?" | tee -a gpurun_out/f.log
python - <<'P'
import json
for f in ('f_bench_c3','f_bench_c3_reference'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, '%.4g'%d['value'], d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('iteration_frac'), 'e2e', d['e2e'].get('value'), d['e2e'].get('seconds'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
P
tail -n 3 gpurun_out/f_bench_c3.err gpurun_out/f_bench_c3_reference.err
