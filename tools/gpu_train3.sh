#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_dgrad.py > gpurun_out/dbg_dgrad.txt 2>&1
