#!/bin/bash
mkdir -p gpurun_out
timeout 100 ./build/mma_microbench 2>&1 | tail -16 > gpurun_out/mma_queue.txt
python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_umma -c 8 -o gpurun_out/conv_v3 python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/ncu8.log 2>&1
