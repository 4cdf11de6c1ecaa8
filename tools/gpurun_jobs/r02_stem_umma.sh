#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -q -m gpu -x > gpurun_out/r02_su_tests.log 2>&1
tail -5 gpurun_out/r02_su_tests.log | cut -c1-300
for v in 1 0; do
  IFCB_STEM_UMMA=$v timeout 300 python tools/run_plan_once.py --batch 1024 --passes 2 --time > gpurun_out/r02_layers_su$v.txt 2>&1; head -2 gpurun_out/r02_layers_su$v.txt; tail -1 gpurun_out/r02_layers_su$v.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_gray3x3_umma -c 1 -f -o gpurun_out/r02_stem_umma python tools/run_plan_once.py --batch 1024 --passes 1 > gpurun_out/r02_ncu_su.log 2>&1
ls -la gpurun_out/r02_stem_umma.ncu-rep
