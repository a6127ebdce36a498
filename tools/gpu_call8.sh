#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/panel_check.py --quick > gpurun_out/c8_panel_parity.log 2>&1; echo "panel parity: $?" | tee gpurun_out/c8.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c8_bench_c4.json 2> gpurun_out/c8_bench_c4.err; echo "bench c4: $?" | tee -a gpurun_out/c8.log
grep -v " ok " gpurun_out/c8_panel_parity.log | tail -n 8; python - <<'P'
import json
for l in open('gpurun_out/c8_bench_c4.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['all_kernels_ms'], d['config']['chains_per_gpu'], d['config']['chunk'], d['config']['groups'], d['viterbi'])
P
tail -3 gpurun_out/c8_bench_c4.err
