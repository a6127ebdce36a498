#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_cuda.py -m gpu -x -q -k "time_sharded or viterbi" > gpurun_out/c11_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c11.log
tail -n 25 gpurun_out/c11_pytest.log
