#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/reference_cache_probe.py > gpurun_out/c21_reference_cache_probe.json 2> gpurun_out/c21_probe.err; echo "probe: $?" | tee gpurun_out/c21.log
cat gpurun_out/c21_reference_cache_probe.json; tail -n 3 gpurun_out/c21_probe.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c21_bench_c3.json 2> gpurun_out/c21_bench_c3.err; echo "bench: $?" | tee -a gpurun_out/c21.log
python - <<'P'
import json
for l in open('gpurun_out/c21_bench_c3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['phases_ms_per_step'], d['roofline']['all_kernels_ms'])
P
