#!/bin/bash
cd /root/repo
L=gpurun_out/r2c31.log
: > $L
run() { echo "== $*" >> $L; timeout 90 python -u tools/check_mma.py "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run --B 2 --H 30 --C 64 --shift 0 --bwd 1 --iters 2
run --bwd 1
run --bwd 1 --B 8 --H 120 --C 128
cat $L
timeout 400 python -m pytest tests/test_attention_gpu.py -q -m gpu -x --timeout 120 2>&1 | tail -5 > gpurun_out/r2c31_tests.log
cat gpurun_out/r2c31_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_mma_bwd -s 1 -c 1 -f -o gpurun_out/r2c31_mma_bwd python tools/check_mma.py --iters 1 --bwd 1 > gpurun_out/r2c31_ncu.log 2>&1
