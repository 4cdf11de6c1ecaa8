#!/bin/bash
# 2-GPU check after the last TRAIN kernel changes: the driver's own command line for N = 2 (RUN + TRAIN block with ddp_check) and the
# two-GPU tests that a 1-GPU box skips
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; head -c 400 gpurun_out/r02_bench_n2.json; echo
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu -k "two_gpu or two_gpus" 2>&1 | tail -2
