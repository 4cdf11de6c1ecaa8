#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -k "pool or forward_backward_vs_oracle or other_families or graph_step" > gpurun_out/r02_pool_tests.log 2>&1; tail -3 gpurun_out/r02_pool_tests.log | cut -c1-300
for a in inception_v3 resnet50; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_pool.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_pool.json')); print('$a','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f mem %.1f GB'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms'],d['mem_gb']))
PY
done
timeout 600 python tools/train_layer_times.py --arch inception_v3 --batch 256 > gpurun_out/r02_tlt_pool.txt 2>&1; head -9 gpurun_out/r02_tlt_pool.txt
