#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build_tools/umma_rate > gpurun_out/umma_rate.log 2>&1; echo "exit $?"; cat gpurun_out/umma_rate.log
