#!/bin/bash
# Round 2, call 2: the wide kernels (17 <= N <= 104) after the register-cap fix: parity + timing, then sanitizers.
mkdir -p gpurun_out
timeout 900 python tests/panel_check.py > gpurun_out/c2_panel_mode1.log 2>&1; echo "panel mode 1 (parity+timing): $?" | tee gpurun_out/c2.log
BHMM_B200_PANEL=1 timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_panel.py > gpurun_out/c2_memcheck_panel.log 2>&1; echo "memcheck: $?" | tee -a gpurun_out/c2.log
BHMM_B200_PANEL=1 timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_panel.py > gpurun_out/c2_racecheck_panel.log 2>&1; echo "racecheck: $?" | tee -a gpurun_out/c2.log
tail -n 30 gpurun_out/c2_panel_mode1.log; tail -n 5 gpurun_out/c2_memcheck_panel.log gpurun_out/c2_racecheck_panel.log
