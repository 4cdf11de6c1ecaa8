#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -k "train_then_run" 2>&1 | tail -40 > gpurun_out/train5.log
