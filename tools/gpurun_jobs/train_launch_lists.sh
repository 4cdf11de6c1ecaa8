#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/train_launches_r50.csv python tools/bench_train.py --arch resnet50 --batch 256 --steps 1 --warmup 1 > gpurun_out/ncu_tr50.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/train_launches_inc.csv python tools/bench_train.py --arch inception_v3 --batch 256 --steps 1 --warmup 1 > gpurun_out/ncu_tinc.log 2>&1
