#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layers_staged.txt 2>&1; head -3 gpurun_out/r02_layers_staged.txt; tail -1 gpurun_out/r02_layers_staged.txt
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "stem or partial or window or default_outfile or run_cli" > gpurun_out/r02_t4.log 2>&1; tail -3 gpurun_out/r02_t4.log
for a in resnet50 inception_v3; do
  for mk in 512 1100 100000; do
  IFCB_TRAIN_EPI_STATS_MINK=$mk timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_mk$mk.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_mk$mk.json')); print('$a mink=$mk','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms']))
PY
  done
done
timeout 600 python tools/bench_cli_run.py --bins-per-gpu 32 > gpurun_out/r02_cli_n1.jsonl 2> gpurun_out/r02_cli_n1.err; cat gpurun_out/r02_cli_n1.jsonl; tail -3 gpurun_out/r02_cli_n1.err
