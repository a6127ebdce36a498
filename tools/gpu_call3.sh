#!/bin/bash
# Round 2, call 3: the panel family as the default: whole GPU suite, C4 at full size, n32 bench, ncu of the wide kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee gpurun_out/c3.log
timeout 600 python tools/c4_full.py > gpurun_out/c3_c4_full.json 2> gpurun_out/c3_c4_full.err; echo "c4 full: $?" | tee -a gpurun_out/c3.log
timeout 300 python bench.py --workload n32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_n32.json 2> gpurun_out/c3_bench_n32.err; echo "bench n32: $?" | tee -a gpurun_out/c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide -c 2 -o gpurun_out/prof_wide13 \
    python tools/c4_full.py --trajectories 1184 --frames 4000 > gpurun_out/c3_ncu_wide.log 2>&1; echo "ncu wide: $?" | tee -a gpurun_out/c3.log
tail -n 5 gpurun_out/c3_pytest_gpu.log; cat gpurun_out/c3_c4_full.json; tail -n 3 gpurun_out/c3_c4_full.err
