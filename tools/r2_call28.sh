#!/bin/bash
cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma_bwd -s 1 -c 1 -f -o gpurun_out/r2c28_mma_bwd python tools/check_mma.py --iters 1 --bwd 1 > gpurun_out/r2c28_ncu.log 2>&1
tail -2 gpurun_out/r2c28_ncu.log
