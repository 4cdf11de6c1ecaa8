#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -x > gpurun_out/r02_train_tests3.log 2>&1
tail -3 gpurun_out/r02_train_tests3.log
for a in resnet50 inception_v3; do
  for v in new whole; do
  if [ $v = whole ]; then export IFCB_BN_REDUCE_WHOLE=1; else unset IFCB_BN_REDUCE_WHOLE; fi
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_bn3_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_bn3_$v.json')); print('$a $v','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms']))
PY
  done
done
