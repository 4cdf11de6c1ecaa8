#!/bin/bash
mkdir -p gpurun_out
for w in 1 3 4; do
IFCB_WGRAD_WAVES=$w timeout 300 python tools/bench_train.py --arch resnet50 --batch 256 --steps 5 --warmup 2 2>/dev/null | grep "^{" > gpurun_out/knob_r50_w$w.json
IFCB_WGRAD_WAVES=$w timeout 300 python tools/bench_train.py --arch inception_v3 --batch 256 --steps 5 --warmup 2 2>/dev/null | grep "^{" > gpurun_out/knob_inc_w$w.json
done
