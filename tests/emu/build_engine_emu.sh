#!/bin/bash
# Builds tests/emu/engine_emu.so: the library's host logic (capi.cu, engine.cu) and its general-N kernels, hostified
# (hostify.py) and compiled for the CPU on top of the warp emulator and the fake CUDA runtime.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
GEN="$HERE/_gen"
mkdir -p "$GEN"
INC="-I${CUDA_HOME:-/usr/local/cuda}/include -I$ROOT/bhmm_b200/csrc -I$HERE"
OBJS=""
for f in capi engine certify team_kernels panel_kernels lane_viterbi scan_kernels frame_kernels sample_kernels transfer; do
    python "$HERE/hostify.py" "$ROOT/bhmm_b200/csrc/$f.cu" > "$GEN/$f.cpp"
    g++ -O1 -std=c++17 -fPIC -w $INC -DPANEL_HOST_EMU=1 -include "$HERE/cuda_fake.h" -c "$GEN/$f.cpp" -o "$GEN/$f.o" &
    OBJS="$OBJS $GEN/$f.o"
done
g++ -O1 -std=c++17 -fPIC -w $INC -c "$HERE/cuda_fake.cpp" -o "$GEN/cuda_fake.o" &
g++ -O1 -std=c++17 -fPIC -w $INC -include "$HERE/cuda_fake.h" -c "$HERE/lane_stubs.cpp" -o "$GEN/lane_stubs.o" &
wait
g++ -shared -o "$HERE/engine_emu.so" $OBJS "$GEN/cuda_fake.o" "$GEN/lane_stubs.o" -lpthread
echo "$HERE/engine_emu.so"
