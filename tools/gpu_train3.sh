#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -4 > gpurun_out/train5.log
timeout 600 python tools/bench_train.py --arch resnet50 --batch 256 --steps 5 --warmup 2 --parts > gpurun_out/bt_r50.json 2> gpurun_out/bt_r50.err
timeout 600 python tools/bench_train.py --arch inception_v3 --batch 256 --steps 5 --warmup 2 --parts > gpurun_out/bt_inc.json 2> gpurun_out/bt_inc.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/train_launches_r50.csv python tools/bench_train.py --arch resnet50 --batch 256 --steps 1 --warmup 1 > gpurun_out/ncu_tr50.log 2>&1
