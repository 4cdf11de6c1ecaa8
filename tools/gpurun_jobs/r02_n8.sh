#!/bin/bash
# 8-GPU evidence: bench line (RUN + TRAIN + in-process DDP check) and file-to-file RUN over bins on disk
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -c 600 gpurun_out/r02_bench_n8.json; tail -3 gpurun_out/r02_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29732 tools/bench_cli_run.py --bins-per-gpu 64 --outfile 'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5' --outfile 'json/{BIN_ID}_class.json' > gpurun_out/r02_cli_n8.jsonl 2> gpurun_out/r02_cli_n8.err
grep bench_cli_run gpurun_out/r02_cli_n8.jsonl | cut -c1-420; tail -3 gpurun_out/r02_cli_n8.err
nproc; free -g | head -2
