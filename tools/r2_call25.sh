#!/bin/bash
cd /root/repo
timeout 300 python tools/check_mma.py --bwd 1 > gpurun_out/r2c25_check.log 2>&1
timeout 300 python tools/check_mma.py --bwd 1 --shift 0 >> gpurun_out/r2c25_check.log 2>&1
timeout 300 python tools/check_mma.py --bwd 1 --B 8 --H 120 --C 128 >> gpurun_out/r2c25_check.log 2>&1
cat gpurun_out/r2c25_check.log
timeout 600 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2c25_tests.log
cat gpurun_out/r2c25_tests.log
