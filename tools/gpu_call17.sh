#!/bin/bash
mkdir -p gpurun_out
for v in "" _vr6; do
  BHMM_B200_LIB=$PWD/bhmm_b200/libbhmm_b200$v.so timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --trajectories 1365 > gpurun_out/c17_c4_1365$v.json 2> gpurun_out/c17_c4$v.err
done
timeout 600 python tests/panel_check.py --quick > gpurun_out/c17_panel_parity.log 2>&1; echo "panel parity: $?" | tee gpurun_out/c17.log
python - <<'P'
import json
for v in ('','_vr6'):
    for l in open('gpurun_out/c17_c4_1365%s.json'%v):
        if l.startswith('{'):
            d=json.loads(l); print(v or 'R4', d['viterbi'])
P
tail -n 3 gpurun_out/c17_c4*.err; grep -a "viterbi N=100\|parity:" gpurun_out/c17_panel_parity.log
