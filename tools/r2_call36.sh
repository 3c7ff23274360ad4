#!/bin/bash
cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_mma_bwd_spec -s 1 -c 1 -f -o gpurun_out/r2c36_spec python tools/check_mma.py --iters 1 --bwd 1 --a 1 --b 1 > gpurun_out/r2c36_ncu.log 2>&1
tail -2 gpurun_out/r2c36_ncu.log
