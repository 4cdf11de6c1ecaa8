#!/bin/bash
# Round-end evidence at the bench default (batch 1024): bench line + ncu launch list + per-launch conv DRAM traffic + TRAIN benches
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/final2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final1.log 2>&1
timeout 600 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/final2_layers.txt 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_umma -s 65 -c 65 --csv --log-file gpurun_out/final2_conv_traffic.csv python tools/run_plan_once.py --batch 1024 --passes 2 > gpurun_out/ncu_final2.log 2>&1
timeout 600 python tools/bench_train.py --arch resnet50 --batch 256 --steps 10 --warmup 3 --parts 2>/dev/null | grep "^{" > gpurun_out/final2_train_r50.json
timeout 600 python tools/bench_train.py --arch inception_v3 --batch 256 --steps 10 --warmup 3 --parts 2>/dev/null | grep "^{" > gpurun_out/final2_train_inc.json
timeout 300 python __graft_entry__.py smoke > gpurun_out/final2_smoke.log 2>&1
