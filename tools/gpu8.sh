#!/bin/bash
mkdir -p gpurun_out
IFCB_CONV_NOPAIR=1 python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/plain8.log 2>&1 &&
IFCB_CONV_NOPAIR=1 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 3 -o gpurun_out/conv_v5 python tools/run_plan_once.py --batch 512 --passes 1 > gpurun_out/ncu8.log 2>&1
