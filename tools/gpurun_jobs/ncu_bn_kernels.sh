#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_train.py --arch resnet50 --batch 256 --steps 1 --warmup 0"
$CMD > gpurun_out/plain_bn.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_apply_kernel -s 0 -c 2 -f -o gpurun_out/bn_apply $CMD > gpurun_out/ncu_bn1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:channel_reduce_kernel -s 0 -c 2 -f -o gpurun_out/bn_reduce0 $CMD > gpurun_out/ncu_bn2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_kernel -s 51 -c 2 -f -o gpurun_out/bn_bwd_apply $CMD > gpurun_out/ncu_bn3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:channel_reduce_kernel -s 104 -c 2 -f -o gpurun_out/bn_reduce1 $CMD > gpurun_out/ncu_bn4.log 2>&1
