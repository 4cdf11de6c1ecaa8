#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/bench_train.py --arch resnet50 --batch 256 --steps 5 --warmup 2 > gpurun_out/bt_r50_n2.json 2> gpurun_out/bt_r50_n2.err
timeout 600 python tools/bench_train.py --arch resnet50 --batch 256 --steps 5 --warmup 2 > gpurun_out/bt_r50_n1.json 2> gpurun_out/bt_r50_n1.err
