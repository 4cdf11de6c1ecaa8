#!/bin/bash
mkdir -p gpurun_out
for v in 128 96 64; do
IFCB_CONV_WINDOW_PAIR_MIN=$v timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layers_wpairmin$v.txt 2>&1; echo "min $v"; grep -E "Conv2d_2b|Conv2d_4a|Mixed_5b.branch5x5_2|Mixed_5b.branch3x3dbl_3|TOTAL" gpurun_out/r02_layers_wpairmin$v.txt
done
