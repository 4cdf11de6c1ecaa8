#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/bench_train.py --arch inception_v3 --batch 256 --steps 10 --warmup 3 > gpurun_out/bt_inc_n2.json 2> gpurun_out/bt_inc_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/bench_train.py --arch resnet50 --batch 256 --steps 10 --warmup 3 > gpurun_out/bt_r50_n2.json 2> gpurun_out/bt_r50_n2.err
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k two_gpu 2>&1 | tail -2 > gpurun_out/ddp_test.log
