#!/bin/bash
# 2-GPU validation of everything the 8-GPU evidence run uses
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 2500 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 tools/bench_cli_run.py --bins-per-gpu 16 > gpurun_out/r02_cli_n2.jsonl 2> gpurun_out/r02_cli_n2.err
grep bench_cli_run gpurun_out/r02_cli_n2.jsonl | cut -c1-400; tail -3 gpurun_out/r02_cli_n2.err
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "two_gpu or deterministic or pool_train" > gpurun_out/r02_n2_tests.log 2>&1; tail -3 gpurun_out/r02_n2_tests.log
