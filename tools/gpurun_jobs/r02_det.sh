#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -k "deterministic or repack or pool_train or stem_im2col or bn_train" > gpurun_out/r02_det_tests.log 2>&1
tail -5 gpurun_out/r02_det_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; tail -c 1500 gpurun_out/r02_bench2.json; tail -3 gpurun_out/r02_bench2.err
