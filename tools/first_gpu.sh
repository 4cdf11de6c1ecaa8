#!/bin/bash
# first GPU shake-down: each stage in its own process (a trapped kernel poisons the CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in "tests/test_preprocess_gpu.py" "tests/test_layers_gpu.py -k probe" "tests/test_layers_gpu.py -k 'pools or stem or head'" "tests/test_layers_gpu.py -k 'conv_bn_relu or fused or many_tiles'" "tests/test_model_gpu.py -s"; do
  name=$(echo "$t" | tr ' /' '__' | tr -d "'")
  echo "=== $t" | tee -a gpurun_out/first.log
  eval timeout 600 python -m pytest $t -q -m gpu -x 2>&1 | tail -40 | tee -a gpurun_out/first.log
done
