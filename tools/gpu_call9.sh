#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/c9_pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee gpurun_out/c9.log
grep -a "denormal_band dev\|passed\|failed\|Error" gpurun_out/c9_pytest_gpu.log | tail -n 12
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/c9_bench_c3.json 2> gpurun_out/c9_bench_c3.err; echo "bench c3: $?" | tee -a gpurun_out/c9.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c9_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-iters 2 > gpurun_out/c9_ncu_launches.log 2>&1; echo "ncu launches: $?" | tee -a gpurun_out/c9.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward_stats_lane -s 2 -c 1 -o gpurun_out/prof_c3_bwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-iters 1 > gpurun_out/c9_ncu_full.log 2>&1; echo "ncu full: $?" | tee -a gpurun_out/c9.log
python - <<'P'
import json
for l in open('gpurun_out/c9_bench_c3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['iteration_frac'], d['e2e']['value'], d['e2e']['seconds'], d['gibbs']['value'], d['cpu_baseline']['value'])
P
tail -3 gpurun_out/c9_bench_c3.err
