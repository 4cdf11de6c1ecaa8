#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s > gpurun_out/r02_train_tests.log 2>&1
grep -E "passed|failed|Error|error|ours" gpurun_out/r02_train_tests.log | tail -8
IFCB_STEM_PACKED=0 timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time 2>&1 | head -8 > gpurun_out/r02_layers_unpacked.txt; timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layers_packed.txt 2>&1; head -3 gpurun_out/r02_layers_unpacked.txt gpurun_out/r02_layers_packed.txt
timeout 900 python -m pytest tests/test_layers_gpu.py -q -m gpu -x > gpurun_out/r02_layer_tests.log 2>&1
tail -3 gpurun_out/r02_layer_tests.log
for a in resnet50 inception_v3; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_epi.json
  IFCB_TRAIN_EPI_STATS=0 timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_noepi.json
  python - <<PY
import json
for t in ('epi','noepi'):
    d=json.load(open('gpurun_out/r02_bt_${a}_%s.json'%t)); print('$a',t,'%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms']))
PY
done
