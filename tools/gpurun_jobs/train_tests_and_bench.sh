#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s 2>&1 | grep -E "teacher|passed|failed|Error|error" | cut -c1-900 | tail -12 > gpurun_out/train5.log
timeout 600 python tools/bench_train.py --arch resnet50 --batch 256 --steps 5 --warmup 2 --parts > gpurun_out/bt_r50.json 2> gpurun_out/bt_r50.err
timeout 600 python tools/bench_train.py --arch inception_v3 --batch 256 --steps 5 --warmup 2 --parts > gpurun_out/bt_inc.json 2> gpurun_out/bt_inc.err
