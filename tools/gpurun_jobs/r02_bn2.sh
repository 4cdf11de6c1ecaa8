#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -x > gpurun_out/r02_train_tests2.log 2>&1
tail -3 gpurun_out/r02_train_tests2.log
for a in resnet50 inception_v3; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_bn2.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_bn2.json')); print('$a','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms']))
PY
done
# one --set full capture of a mid-size pair conv (ResNet-50 layer3 1x1) to see what binds it
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_pair_kernel -s 30 -c 3 -f -o gpurun_out/r02_pair_mid python tools/bench_train.py --arch resnet50 --batch 256 --steps 1 --warmup 0 > gpurun_out/r02_ncu_pair.log 2>&1
ls -la gpurun_out/r02_pair_mid.ncu-rep
