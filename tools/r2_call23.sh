#!/bin/bash
cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma_fwd -s 2 -c 1 -f -o gpurun_out/r2c23_mma_fwd python tools/check_mma.py --iters 1 > gpurun_out/r2c23_ncu.log 2>&1
tail -3 gpurun_out/r2c23_ncu.log
