#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu -s -k "other_families" > gpurun_out/r02_fam_tests.log 2>&1
grep -E "loss|passed|failed|Error|error:|assert" gpurun_out/r02_fam_tests.log | cut -c1-300 | tail -40
