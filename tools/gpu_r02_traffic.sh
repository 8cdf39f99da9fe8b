#!/bin/bash
mkdir -p gpurun_out
python -c "
import torch
p = torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size, 'persisting max', getattr(p, 'persisting_l2_cache_max_size', None), 'window max', getattr(p, 'access_policy_max_window_size', None))"
timeout 200 python tools/bench_i8.py > gpurun_out/bench_i8.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/bench_i8.log
timeout 240 ncu --clock-control none --target-processes application-only --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_score_i8 -s 2 -c 2 --csv --log-file gpurun_out/i8_traffic.csv python tools/profile_score.py 5 > gpurun_out/ncu_i8_traffic.log 2>&1; echo "ncu exit $?"; grep -v "^==" gpurun_out/i8_traffic.csv | cut -d, -f5,13- | tail -10
