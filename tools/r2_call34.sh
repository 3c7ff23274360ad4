#!/bin/bash
cd /root/repo
L=gpurun_out/r2c34.log
: > $L
run() { echo "== $*" >> $L; timeout 90 python -u tools/check_mma.py "$@" 2>&1 | grep -E "impl|dtable|dq |dv |dvpad|dscale|out  |Error|error" >> $L; echo "rc=$?" >> $L; }
run --bwd 1
run --bwd 1 --shift 0
run --bwd 1 --B 8 --H 120 --C 128
run --bwd 1 --ws 6 --shift 3 --B 48 --H 15 --C 1024
cat $L
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 > gpurun_out/r2c34_tests.log
cat gpurun_out/r2c34_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err
grep -o '"value": [0-9.]*' gpurun_out/r2c34_bench.json | head -2
grep breakdown gpurun_out/r2c34_bench.err
