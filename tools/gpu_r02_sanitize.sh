#!/bin/bash
# compute-sanitizer memcheck over the tcgen05 scoring kernel (small shapes) and the stepped Eagle / set-PE / stack tests
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_score_i8.py -x -q -k "192 or 520 or refit" > gpurun_out/sanitize_i8_r02.log 2>&1; echo "memcheck i8 exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_i8_r02.log | head -8
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stepped_eagle.py tests/test_gpu_transfer.py -x -q -k "set_pe_score or stack_score or out_of_order" > gpurun_out/sanitize_step_r02.log 2>&1; echo "memcheck step exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_step_r02.log | head -8
