#!/bin/bash
# First GPU call of the next round, in one gpurun (about 12-15 minutes of box time):
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# 1. the default GPU suite (must stay green), 2. parity + timing of the opt-in panel family (never run on hardware before:
# every step under its own timeout, a hang must not take the box), 3. bench lines at N = 32 with and without it,
# 4. ncu launch list + one full capture of the panel statistics kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu: $?" | tee -a gpurun_out/first_call.log
timeout 420 python tests/panel_check.py --quick > gpurun_out/panel_parity_mode1.log 2>&1; echo "panel parity mode 1: $?" | tee -a gpurun_out/first_call.log
timeout 300 python tests/panel_check.py --quick --mode=2 > gpurun_out/panel_parity_mode2.log 2>&1; echo "panel parity mode 2: $?" | tee -a gpurun_out/first_call.log
timeout 600 python tests/panel_check.py > gpurun_out/panel_timing_mode1.log 2>&1; echo "panel timing mode 1: $?" | tee -a gpurun_out/first_call.log
timeout 300 python tests/panel_check.py --mode=2 > gpurun_out/panel_timing_mode2.log 2>&1; echo "panel timing mode 2: $?" | tee -a gpurun_out/first_call.log
for mode in 0 1 2; do
    BHMM_B200_PANEL=$mode timeout 300 python bench.py --workload n32 --steps 5 --warmup 3 --no-cpu-baseline \
        > gpurun_out/bench_n32_panel$mode.json 2> gpurun_out/bench_n32_panel$mode.err
    echo "bench n32 panel=$mode: $?" | tee -a gpurun_out/first_call.log
done
for mode in 0 1; do
    BHMM_B200_PANEL=$mode timeout 240 python tools/c5_viterbi.py --frames 2e7 > gpurun_out/c5_viterbi_2e7_panel$mode.json 2>> gpurun_out/first_call.log
done
BHMM_B200_PANEL=1 timeout 300 python tools/c5_viterbi.py --frames 1e9 > gpurun_out/c5_viterbi_1e9_panel1.json 2>> gpurun_out/first_call.log
BHMM_B200_PANEL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_n32_panel1.csv python bench.py --workload n32 --trajectories 64 --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/ncu_n32_panel1.log 2>&1
BHMM_B200_PANEL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel32 -c 2 \
    -o gpurun_out/prof_panel32 python bench.py --workload n32 --trajectories 64 --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/ncu_full_panel32.log 2>&1
tail -n 40 gpurun_out/panel_timing_mode1.log gpurun_out/panel_parity_mode1.log gpurun_out/first_call.log
