#!/bin/bash
mkdir -p gpurun_out
echo "== base_offset mode 1 (default)" > gpurun_out/tests4.log
timeout 900 python -m pytest tests/test_layers_gpu.py -q -m gpu 2>&1 | tail -25 >> gpurun_out/tests4.log
echo "== base_offset mode 0" >> gpurun_out/tests4.log
IFCB_WINDOW_BASE_OFFSET=0 timeout 900 python -m pytest tests/test_layers_gpu.py -q -m gpu -k "window or bn_relu" 2>&1 | tail -25 >> gpurun_out/tests4.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_preprocess_gpu.py -q -m gpu -s 2>&1 | tail -30 >> gpurun_out/tests4.log
python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers4.txt 2>&1
