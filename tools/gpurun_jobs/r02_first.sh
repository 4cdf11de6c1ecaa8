#!/bin/bash
# round 2, first pass: new tests (bench-config parity, output window, default .h5 CLI, short last batch), bench line with the TRAIN block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu -x -s -k "benchmarked or window or default_outfile or whole_model or partial" > gpurun_out/r02_t1.log 2>&1
tail -5 gpurun_out/r02_t1.log
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s -k "short_last or train_then_run or graph_step" > gpurun_out/r02_t2.log 2>&1
tail -5 gpurun_out/r02_t2.log
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -c 3000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
