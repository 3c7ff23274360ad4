#!/bin/bash
cd /root/repo
timeout 300 python tools/check_mma.py > gpurun_out/r2c24_check.log 2>&1
timeout 300 python tools/check_mma.py --shift 0 >> gpurun_out/r2c24_check.log 2>&1
cat gpurun_out/r2c24_check.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma_fwd -s 2 -c 1 -f -o gpurun_out/r2c24_mma_fwd python tools/check_mma.py --iters 1 > gpurun_out/r2c24_ncu.log 2>&1
