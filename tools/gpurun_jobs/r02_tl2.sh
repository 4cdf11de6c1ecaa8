#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "conv_wgrad or forward_backward_vs_oracle" 2>&1 | tail -1
for a in inception_v3 resnet50; do
  timeout 600 python tools/bench_train.py --arch $a --batch 256 --steps 20 --warmup 5 --graph --parts 2>/dev/null | grep "^{" > gpurun_out/r02_bt_${a}_wglean3.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bt_${a}_wglean3.json')); print('$a','%.1f img/s %.2f ms fwd %.2f bwd %.2f opt %.2f mem %.1f GB'%(d['value'],d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['adam_repack_ms'],d['mem_gb']))
PY
done
for a in resnet50 inception_v3; do
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_train_launches_final_$a.csv python tools/bench_train.py --arch $a --batch 256 --steps 1 --warmup 1 > gpurun_out/r02_ncu_tf_$a.log 2>&1
done
