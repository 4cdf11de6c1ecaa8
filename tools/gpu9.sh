#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py -q -m gpu -x 2>&1 | tail -4 > gpurun_out/tests9.log
timeout 600 python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers9_w16.txt 2>&1
cp ifcb_classifier_b200/libifcb_b200.so /tmp/w16.so; cp build/libifcb_b200_w12.so ifcb_classifier_b200/libifcb_b200.so
timeout 600 python tools/run_plan_once.py --batch 512 --passes 2 --time > gpurun_out/layers9_w12.txt 2>&1
cp /tmp/w16.so ifcb_classifier_b200/libifcb_b200.so
