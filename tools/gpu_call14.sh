#!/bin/bash
mkdir -p gpurun_out
python tools/scan_debug.py team > gpurun_out/c14_scan_debug.log 2>&1
grep -a "ERR\|ok\|^-1" gpurun_out/c14_scan_debug.log | head -n 45
