#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -s -k "conv_wgrad or forward_backward_vs_oracle or deterministic or graph_step" > gpurun_out/r02_wgwin_tests.log 2>&1
grep -E "passed|failed|Error|error:|assert" gpurun_out/r02_wgwin_tests.log | cut -c1-300 | tail -6
for a in inception_v3 resnet50; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_wglean2.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_wglean2.json')); print('$a','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f mem %.1f GB'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms'],d['mem_gb']))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_wgrad -c 400 --csv --log-file gpurun_out/r02_wgwin_launches_lean.csv python tools/bench_train.py --arch inception_v3 --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_wgwin_lean.log 2>&1
