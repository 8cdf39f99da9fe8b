#!/bin/bash
# round-2 ncu evidence: launch lists + one `--set full` capture per kernel of interest.  Every step is bounded
# (a wedged profiler otherwise burns the whole call) and the script stops at the first profiler failure.
mkdir -p gpurun_out
NCU="ncu --clock-control none --target-processes application-only"
step() { name=$1; shift; timeout 240 "$@" > gpurun_out/ncu_${name}.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -2 gpurun_out/ncu_${name}.log | cut -c1-200; if [ $rc -ne 0 ]; then echo "profiler step $name failed: stopping"; exit 1; fi; }
# 1. k_score, full set (the kernel the roofline is quoted on)
step score_r02 $NCU --set full --import-source on -k regex:k_score -s 2 -c 1 -f -o gpurun_out/score_r02 python tools/profile_score.py 4
# 2. the bench step: launch list (cold-cache, serialised: compare shares)
step bench_r02 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-suggest --no-cpu
# 3. the fit / NLL path: launch list
step fit_r02 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_fit_r02.csv python tools/profile_fit.py
# 4. the dataflow factorisation kernel, full set (fit at N=1000)
step dataflow_r02 $NCU --set full --import-source on -k regex:k_chol_dataflow -s 2 -c 1 -f -o gpurun_out/dataflow_r02 python tools/profile_fit.py
# 5. the tcgen05 scoring kernel, full set
VZGP_SCORE_I8=1 step score_i8_r02 $NCU --set full --import-source on -k regex:k_score_i8 -s 2 -c 1 -f -o gpurun_out/score_i8_r02 python tools/profile_score.py 4
# 6. small-pool W kernel and the cooperative Eagle grid kernel
step var_small_r02 $NCU --set full --import-source on -k regex:k_var_small -s 2 -c 1 -f -o gpurun_out/var_small_r02 python tools/profile_small_score.py
step eagle_grid_r02 $NCU --set full --import-source on -k regex:k_eagle_grid -s 1 -c 1 -f -o gpurun_out/eagle_grid_r02 python tools/profile_eagle_grid.py
ls -la gpurun_out/*.ncu-rep
