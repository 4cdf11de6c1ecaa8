#!/bin/bash
# Round-end evidence for the RUN path: GPU test suite, bench line, ncu launch list, DRAM traffic of every conv launch
# of one batch (metrics only: a --set full report of 65 launches exceeds the 64 MiB copy-back limit), --set full of
# three representative conv launches and of the preprocess kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/final_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final1.log 2>&1
timeout 600 python tools/run_plan_once.py --batch 512 --passes 2 > gpurun_out/plain_final2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_umma -s 65 -c 65 --csv --log-file gpurun_out/final_conv_traffic.csv python tools/run_plan_once.py --batch 512 --passes 2 > gpurun_out/ncu_final2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_umma -s 67 -c 1 -f -o gpurun_out/final_conv_2b python tools/run_plan_once.py --batch 512 --passes 2 > gpurun_out/ncu_final3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_umma -s 110 -c 1 -f -o gpurun_out/final_conv_6e python tools/run_plan_once.py --batch 512 --passes 2 > gpurun_out/ncu_final4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:preprocess_kernel -s 4 -c 1 -f -o gpurun_out/final_preprocess python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final5.log 2>&1
ls -la gpurun_out > gpurun_out/final_ls.txt
