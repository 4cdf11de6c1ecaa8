#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 66 -c 1 -f -o gpurun_out/r02_wgrad_2a python tools/bench_train.py --arch inception_v3 --batch 256 --steps 1 --warmup 0 > gpurun_out/r02_ncu_wg2a.log 2>&1
ls -la gpurun_out/r02_wgrad_2a.ncu-rep
