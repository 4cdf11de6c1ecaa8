#!/bin/bash
mkdir -p gpurun_out
for a in inception_v3 resnet50; do
for v in 1 2; do
  IFCB_WGRAD_WAVES=$v timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_waves$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_waves$v.json')); print('$a waves=$v','%.1f img/s %.2f ms fwd %.2f bwd %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms']))
PY
done
done
