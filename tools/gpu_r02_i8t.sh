#!/bin/bash
mkdir -p gpurun_out
for v in i8timing i8t_STORES i8t_MATERN; do
VZGP_LIB=$PWD/vizier_b200/_lib/libvzgp_$v.so timeout 200 python tools/i8_timing.py > gpurun_out/i8_timing_$v.log 2>&1; echo "$v exit $?"; tail -1 gpurun_out/i8_timing_$v.log
done
