#!/bin/bash
mkdir -p gpurun_out
python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_umma -c 3 -o gpurun_out/conv_v4 python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/ncu8.log 2>&1
