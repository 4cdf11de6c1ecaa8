#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dbg_poolafter.py > gpurun_out/dbg_pa.txt 2>&1
