#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c25.log
tail -n 5 gpurun_out/c25_pytest.log
timeout 300 python tools/e2e_breakdown.py > gpurun_out/c25_e2e_breakdown.log 2>&1; tail -n 3 gpurun_out/c25_e2e_breakdown.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c25_bench_c3.json 2> gpurun_out/c25_bench_c3.err; echo "bench: $?" | tee -a gpurun_out/c25.log
python - <<'P'
import json
for f in ('c25_bench_c3',):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['all_kernels_ms'], d['config']['certification'], 'e2e %.4g in %.3f s'%(d['e2e']['value'], d['e2e']['seconds']))
P
tail -n 3 gpurun_out/c25_bench_c3.err
