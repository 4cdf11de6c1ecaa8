#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s -k "oracle and inception" 2>&1 | grep -E "teacher|end to end|Error|passed|failed" | cut -c1-1200 > gpurun_out/train6.log
