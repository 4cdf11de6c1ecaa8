#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "pool_train or weight_refresh or train_then_run or forward_backward or short_last" > gpurun_out/r02_misc_tests.log 2>&1; tail -4 gpurun_out/r02_misc_tests.log
timeout 600 python tools/bench_cli_run.py --bins-per-gpu 64 --outfile 'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5' --outfile 'json/{BIN_ID}_class.json' > gpurun_out/r02_cli_n1_64.jsonl 2> gpurun_out/r02_cli_n1_64.err; grep bench_cli gpurun_out/r02_cli_n1_64.jsonl | cut -c1-330
for a in resnet50 inception_v3; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_v4.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_v4.json')); print('$a','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms']))
PY
done
