#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/tests3.log
python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers3.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 65 -c 36 -o gpurun_out/conv_full python tools/run_plan_once.py --batch 512 --passes 2 > gpurun_out/ncu3.log 2>&1
