#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee gpurun_out/c6.log
timeout 300 python tools/e2e_breakdown.py > gpurun_out/c6_e2e_breakdown.log 2>&1; echo "e2e breakdown: $?" | tee -a gpurun_out/c6.log
timeout 300 python tools/benchmark_hidden.py > gpurun_out/c6_benchmark_hidden.log 2>&1; echo "benchmark_hidden: $?" | tee -a gpurun_out/c6.log
tail -n 15 gpurun_out/c6_pytest_gpu.log; cat gpurun_out/c6_e2e_breakdown.log; tail -n 20 gpurun_out/c6_benchmark_hidden.log
