#!/bin/bash
mkdir -p gpurun_out
for b in 1024 2048; do
timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
done
