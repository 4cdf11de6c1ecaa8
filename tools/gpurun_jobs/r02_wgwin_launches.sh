#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
IFCB_WGRAD_WINDOW=$v timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_wgrad -c 400 --csv --log-file gpurun_out/r02_wgwin_launches_$v.csv python tools/bench_train.py --arch inception_v3 --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_wgwin_$v.log 2>&1
done
ls -la gpurun_out/r02_wgwin_launches_*.csv
