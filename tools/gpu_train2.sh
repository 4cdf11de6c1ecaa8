#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu -s -k "oracle" 2>&1 | grep -E "teacher|end to end|passed|failed|ours|Error" | cut -c1-1500 > gpurun_out/train3.log
