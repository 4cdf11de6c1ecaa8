#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess_gpu.py -q -m gpu -x > gpurun_out/r02_pre_tests.log 2>&1
tail -15 gpurun_out/r02_pre_tests.log
timeout 300 python tools/bench_preprocess.py > gpurun_out/r02_pre_bench.txt 2>&1
cat gpurun_out/r02_pre_bench.txt
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -s -k "benchmarked" > gpurun_out/r02_t3.log 2>&1
tail -4 gpurun_out/r02_t3.log
