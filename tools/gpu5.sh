#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -q -m gpu 2>&1 | tail -5 > gpurun_out/tests5.log
timeout 1200 python tools/bench_layer.py > gpurun_out/layer_matrix.txt 2>&1
