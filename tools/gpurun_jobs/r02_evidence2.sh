#!/bin/bash
# Round-2 evidence refresh on one B200 after the tensor-core stem / wgrad WINDOW / sibling-fusion changes: full GPU test suite, bench
# lines (fp16 default incl. the TRAIN block, bf16), smoke, per-layer table, ncu launch list of the bench command, per-launch conv DRAM
# traffic, TRAIN launch lists.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; head -c 300 gpurun_out/r02_bench_final.json; echo
timeout 600 python bench.py --dtype bf16 --no-train --no-cpu-baseline > gpurun_out/r02_bench_final_bf16.json 2> gpurun_out/r02_bench_final_bf16.err; head -c 200 gpurun_out/r02_bench_final_bf16.json; echo
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1; tail -1 gpurun_out/r02_smoke_final.log
timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layer_events_final_b1024.txt 2>&1; tail -1 gpurun_out/r02_layer_events_final_b1024.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --bins 4 --no-train --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_umma -s 65 -c 65 --csv --log-file gpurun_out/r02_conv_traffic.csv python tools/run_plan_once.py --batch 1024 --passes 2 > gpurun_out/r02_ncu_c.log 2>&1
for a in resnet50 inception_v3; do
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_train_launches_final_$a.csv python tools/bench_train.py --arch $a --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_tf_$a.log 2>&1
done
ls -la gpurun_out/r02_train_launches_final_*.csv
