#!/bin/bash
# compute-sanitizer passes over the small-size parity tests (memcheck, then racecheck on the score kernel).
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel_matrix or cross_kernel or cholesky or fit_factor or score_with_aux or score_edge or nll_grad or topk or eagle or random_search or posterior or score_topk or small_pool or ensemble_score or pack or nll_grad_small or ensemble_eagle" \
  > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck.log; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "score_edge or cholesky_and_inverse or nll_grad or small_pool or pack or eagle_run_matches or pe_acquisition" \
  > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/sanitize_racecheck.log; grep -c "hazard" gpurun_out/sanitize_racecheck.log
