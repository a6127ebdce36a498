#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/panel_check.py --quick > gpurun_out/c16_panel_parity.log 2>&1; echo "panel parity: $?" | tee gpurun_out/c16.log
grep -a "viterbi\|FAIL\|parity:" gpurun_out/c16_panel_parity.log | tail -n 8
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c16_bench_c4.json 2> gpurun_out/c16_bench_c4.err; echo "bench c4: $?" | tee -a gpurun_out/c16.log
BHMM_B200_VITERBI_REGS=1 timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --trajectories 1184 > gpurun_out/c16_bench_c4_regs.json 2> gpurun_out/c16_bench_c4_regs.err
timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --trajectories 1184 > gpurun_out/c16_bench_c4_multi.json 2> gpurun_out/c16_bench_c4_multi.err
python - <<'P'
import json
for f in ('c16_bench_c4','c16_bench_c4_regs','c16_bench_c4_multi'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], d['viterbi'])
P
tail -n 3 gpurun_out/c16_bench_c4.err
