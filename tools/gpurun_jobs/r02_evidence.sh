#!/bin/bash
# Round-2 evidence on one B200: full GPU test suite, bench lines (fp16 default + bf16), ncu launch list of the bench command, per-layer
# table, per-launch conv DRAM traffic, ncu --set full of the preprocess kernel in its three output modes and of the TRAIN streaming kernels,
# TRAIN launch lists, smoke.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; head -c 300 gpurun_out/r02_bench_final.json; echo
timeout 600 python bench.py --dtype bf16 --no-train --no-cpu-baseline > gpurun_out/r02_bench_final_bf16.json 2> gpurun_out/r02_bench_final_bf16.err; head -c 200 gpurun_out/r02_bench_final_bf16.json; echo
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1; tail -1 gpurun_out/r02_smoke_final.log
timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layer_events_final_b1024.txt 2>&1; tail -1 gpurun_out/r02_layer_events_final_b1024.txt
timeout 300 python tools/bench_preprocess.py > gpurun_out/r02_preprocess_timing.txt 2>&1
# launch list of the bench command (RUN part) -- the same command first runs without ncu (above)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --bins 4 --no-train --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
# per-launch conv DRAM traffic / tensor-pipe activity of one 1024-ROI batch (second pass: launches 65..129)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_umma -s 65 -c 65 --csv --log-file gpurun_out/r02_conv_traffic.csv python tools/run_plan_once.py --batch 1024 --passes 2 > gpurun_out/r02_ncu_c.log 2>&1
# preprocess kernel, --set full, three output modes (tools/bench_preprocess.py launches u8 x2 bounds, f32 x2, bf16 x2; 3 warm-ups + 20 timed each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 3 -c 1 -f -o gpurun_out/r02_pre_u8 python tools/bench_preprocess.py > gpurun_out/r02_ncu_p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 49 -c 1 -f -o gpurun_out/r02_pre_f32 python tools/bench_preprocess.py > gpurun_out/r02_ncu_p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 95 -c 1 -f -o gpurun_out/r02_pre_bf16 python tools/bench_preprocess.py > gpurun_out/r02_ncu_p3.log 2>&1
# TRAIN: launch lists of one step
for a in resnet50 inception_v3; do
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_train_launches_final_$a.csv python tools/bench_train.py --arch $a --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_tf_$a.log 2>&1
done
# TRAIN streaming kernels, --set full on ResNet-50's largest layers
timeout 600 ncu --set full --clock-control none -k regex:"channel_reduce_kernel|bn_apply_kernel|bn_bwd_apply_kernel|stem_im2col_tiled|conv_repack_batch|maxpool3" -s 0 -c 12 -f -o gpurun_out/r02_train_stream python tools/bench_train.py --arch resnet50 --batch 256 --steps 1 --warmup 0 > gpurun_out/r02_ncu_ts.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
