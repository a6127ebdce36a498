#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c29_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c29.log
tail -n 5 gpurun_out/c29_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c29_bench_c3.json 2> gpurun_out/c29_bench_c3.err; echo "bench: $?" | tee -a gpurun_out/c29.log
python - <<'P'
import json
for f in ('c29_bench_c3',):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['all_kernels_ms'], 'e2e %.4g in %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), d['e2e']['phases_s'], 'gibbs %.4g' % d['gibbs']['value'])
P
tail -n 3 gpurun_out/c29_bench_c3.err
