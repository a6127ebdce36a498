#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/panel_check.py > gpurun_out/c7_panel_mode1.log 2>&1; echo "panel mode 1 (parity+timing): $?" | tee gpurun_out/c7.log
BHMM_B200_PANEL=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_panel.py > gpurun_out/c7_racecheck_panel.log 2>&1; echo "racecheck: $?" | tee -a gpurun_out/c7.log
BHMM_B200_PANEL=1 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_panel.py > gpurun_out/c7_memcheck_panel.log 2>&1; echo "memcheck: $?" | tee -a gpurun_out/c7.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c7_bench_c4.json 2> gpurun_out/c7_bench_c4.err; echo "bench c4: $?" | tee -a gpurun_out/c7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide2 -c 2 -o gpurun_out/prof_wide2_13 \
    python tools/c4_full.py --trajectories 2368 --frames 4000 > gpurun_out/c7_ncu_wide2.log 2>&1; echo "ncu wide2: $?" | tee -a gpurun_out/c7.log
grep -v " ok " gpurun_out/c7_panel_mode1.log | tail -n 20; tail -n 4 gpurun_out/c7_racecheck_panel.log gpurun_out/c7_memcheck_panel.log; cat gpurun_out/c7_bench_c4.json | cut -c1-1500; tail -3 gpurun_out/c7_bench_c4.err
