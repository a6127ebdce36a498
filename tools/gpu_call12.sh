#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/c12_bench_c5.json 2> gpurun_out/c12_bench_c5.err; echo "bench c5: $?" | tee gpurun_out/c12.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 bench.py --workload c5 --frames 1e8 --steps 2 --warmup 1 > gpurun_out/c12_bench_c5_tr.json 2> gpurun_out/c12_bench_c5_tr.err; echo "bench c5 torchrun: $?" | tee -a gpurun_out/c12.log
python - <<'P'
import json
for f in ('gpurun_out/c12_bench_c5.json','gpurun_out/c12_bench_c5_tr.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['all_kernels_ms'], d['config']['certification'], d['viterbi'])
P
tail -5 gpurun_out/c12_bench_c5.err gpurun_out/c12_bench_c5_tr.err
