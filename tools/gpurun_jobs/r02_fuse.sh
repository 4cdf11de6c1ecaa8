#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -s -k "forward_backward_vs_oracle or graph_step or other_families or train_steps_follow" > gpurun_out/r02_fuse_tests.log 2>&1
grep -E "teacher|passed|failed|Error|error:|assert" gpurun_out/r02_fuse_tests.log | cut -c1-400 | tail -12
for v in 1 0; do
  IFCB_TRAIN_FUSE_SIBLINGS=$v timeout 600 python tools/bench_train.py --arch inception_v3 --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_inc_fuse$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_inc_fuse$v.json')); print('inception fuse=$v','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f mem %.1f GB'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms'],d['mem_gb']))
PY
done
