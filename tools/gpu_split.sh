#!/bin/bash
# split-walk latency mode: parity + single-frame times for several hand-over thresholds (C2, then C4)
set -u
mkdir -p gpurun_out
timeout 300 python tools/split_check.py C2 "$@" > gpurun_out/split_check_c2.txt 2> gpurun_out/split_check.err || { echo "SPLIT CHECK C2 HUNG/FAILED"; tail -5 gpurun_out/split_check.err; }
cat gpurun_out/split_check_c2.txt
