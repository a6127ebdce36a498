#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c23.log
tail -n 3 gpurun_out/c23_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c23_bench_c3.json 2> gpurun_out/c23_bench_c3.err; echo "bench: $?" | tee -a gpurun_out/c23.log
BHMM_B200_OPTIMISTIC=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c23_bench_c3_pess.json 2> gpurun_out/c23_bench_c3_pess.err
python - <<'P'
import json
for f in ('c23_bench_c3','c23_bench_c3_pess'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], d['config']['phases_ms_per_step']['rank0'], d['roofline']['all_kernels_ms'], d['config']['certification'], d['gpu_launches'])
P
tail -n 3 gpurun_out/c23_bench_c3.err
