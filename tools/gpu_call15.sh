#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c15.log
grep -a "RuntimeError\|passed\|failed\|AssertionError\|assert \|Error" gpurun_out/c15_pytest.log | tail -n 20
