#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests6.log
timeout 600 python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers6.txt 2>&1
timeout 900 python tools/bench_layer.py > gpurun_out/layer_matrix6.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err
