#!/bin/bash
cd /root/repo
for k in 1; do timeout 60 python -u tools/check_mma.py --bwd 1 --a 1 --b 3 2>&1 | grep -E "impl|out  |lse|dq |dtable"; done
timeout 60 python -u tools/check_mma.py --bwd 0 --a 1 --b 3 --shift 0 2>&1 | grep -E "impl|out  "
timeout 60 python -u tools/check_mma.py --bwd 0 --a 1 --b 3 --B 8 --H 120 --C 128 2>&1 | grep -E "impl|out  "
timeout 60 python -u tools/check_mma.py --bwd 0 --a 1 --b 3 --ws 6 --shift 3 --B 48 --H 15 --C 1024 2>&1 | grep -E "impl|out  "
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x --timeout 120 2>&1 | tail -15 > gpurun_out/r2c39_tests.log
cat gpurun_out/r2c39_tests.log
