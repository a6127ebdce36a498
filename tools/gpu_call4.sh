#!/bin/bash
# Round 2, call 4: new GPU tests (reference estimators on 'cuda', device M-step), restructured bench, C4 with the round planner
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee gpurun_out/c4.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench_c3.json 2> gpurun_out/c4_bench_c3.err; echo "bench c3: $?" | tee -a gpurun_out/c4.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c4_bench_c3_reference.json 2> gpurun_out/c4_bench_ref.err; echo "bench ref: $?" | tee -a gpurun_out/c4.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c4_bench_c4.json 2> gpurun_out/c4_bench_c4.err; echo "bench c4: $?" | tee -a gpurun_out/c4.log
timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/c4_bench_c5.json 2> gpurun_out/c4_bench_c5.err; echo "bench c5: $?" | tee -a gpurun_out/c4.log
tail -n 15 gpurun_out/c4_pytest_gpu.log; tail -n 5 gpurun_out/c4_bench_c3.err gpurun_out/c4_bench_c4.err gpurun_out/c4_bench_c5.err gpurun_out/c4_bench_ref.err
