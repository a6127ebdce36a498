#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee gpurun_out/c5.log
timeout 300 python tools/e2e_breakdown.py > gpurun_out/c5_e2e_breakdown.log 2>&1; echo "e2e breakdown: $?" | tee -a gpurun_out/c5.log
tail -n 15 gpurun_out/c5_pytest_gpu.log; cat gpurun_out/c5_e2e_breakdown.log
