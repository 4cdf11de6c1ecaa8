#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -s 2>&1 | tail -40 > gpurun_out/tests2.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
python bench.py --steps 5 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rois 512 > gpurun_out/ncu.log 2>&1
