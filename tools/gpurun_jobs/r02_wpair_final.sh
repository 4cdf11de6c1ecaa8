#!/bin/bash
# CTA-pair WINDOW convs: full GPU suite, bench line (RUN + TRAIN block), per-layer table
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu ) > gpurun_out/r02_gpu_tests_final.log 2>&1; grep -E "passed|failed|^FAILED" gpurun_out/r02_gpu_tests_final.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; head -c 250 gpurun_out/r02_bench_final.json; echo
timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layer_events_final_b1024.txt 2>&1; tail -1 gpurun_out/r02_layer_events_final_b1024.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
