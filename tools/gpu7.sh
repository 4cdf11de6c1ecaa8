#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/tests7.log
timeout 600 python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers7.txt 2>&1
timeout 900 python tools/bench_layer.py > gpurun_out/layer_matrix7.txt 2>&1
