#!/bin/bash
mkdir -p gpurun_out
for a in resnet50 inception_v3; do
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip-before-match 0 -c 6000 --csv --log-file gpurun_out/r02_train_launches_$a.csv python tools/bench_train.py --arch $a --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_t_$a.log 2>&1
done
ls -la gpurun_out/r02_train_launches_*.csv
