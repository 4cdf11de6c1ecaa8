#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu -s -k "dgrad or oracle" 2>&1 | grep -E "vs |passed|failed|ours|Error" | cut -c1-400 > gpurun_out/train3.log
