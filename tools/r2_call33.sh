#!/bin/bash
cd /root/repo
L=gpurun_out/r2c33.log
: > $L
run() { echo "== $*" >> $L; timeout 90 python -u tools/check_mma.py "$@" 2>&1 | grep -E "impl|dtable|dq |out  |Error|error" >> $L; echo "rc=$?" >> $L; }
run --bwd 1 --ws 6 --shift 3 --B 48 --H 15 --C 1024
run --bwd 1 --ws 6 --shift 0 --B 48 --H 15 --C 1024
run --bwd 1 --ws 8 --shift 4 --B 16 --H 60 --C 256
run --bwd 1 --ws 7 --shift 3 --B 16 --H 60 --C 256
run --bwd 1 --ws 4 --shift 2 --B 16 --H 60 --C 256
cat $L
timeout 400 python -m pytest tests/test_attention_gpu.py -q -m gpu -x --timeout 120 2>&1 | tail -5 > gpurun_out/r2c33_tests.log
cat gpurun_out/r2c33_tests.log
