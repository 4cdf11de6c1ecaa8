#!/bin/bash
# after the CTA-pair WINDOW convs: bf16 bench line, ncu launch list of the bench command, per-launch conv DRAM traffic
mkdir -p gpurun_out
timeout 600 python bench.py --dtype bf16 --no-train --no-cpu-baseline > gpurun_out/r02_bench_final_bf16.json 2> gpurun_out/r02_bench_final_bf16.err; head -c 200 gpurun_out/r02_bench_final_bf16.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --bins 4 --no-train --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_umma -s 65 -c 65 --csv --log-file gpurun_out/r02_conv_traffic.csv python tools/run_plan_once.py --batch 1024 --passes 2 > gpurun_out/r02_ncu_c.log 2>&1
ls -la gpurun_out/r02_ncu_launches_bench.csv gpurun_out/r02_conv_traffic.csv
