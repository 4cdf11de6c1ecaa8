#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_layer_times.py --arch inception_v3 --batch 256 > gpurun_out/r02_train_layer_times_inception.txt 2>&1; tail -75 gpurun_out/r02_train_layer_times_inception.txt | head -40
timeout 600 python tools/train_layer_times.py --arch resnet50 --batch 256 > gpurun_out/r02_train_layer_times_resnet50.txt 2>&1; head -20 gpurun_out/r02_train_layer_times_resnet50.txt
