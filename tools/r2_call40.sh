#!/bin/bash
cd /root/repo
L=gpurun_out/r2c40.log
: > $L
run() { echo "== $*" >> $L; timeout 60 python -u tools/check_mma.py "$@" 2>&1 | grep -E "impl|dtable|dq |dv |dvpad|dscale|Error|error|column" >> $L; echo "rc=${PIPESTATUS[0]}" >> $L; }
run --bwd 1 --a 1 --b 4 --B 2 --H 30 --C 64 --shift 6 --iters 2
run --bwd 1 --a 1 --b 4
cat $L
if grep -q "rc=124" $L; then echo HANG; exit 0; fi
timeout 900 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -4 > gpurun_out/r2c40_tests.log
cat gpurun_out/r2c40_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r2c40_bench.json 2> gpurun_out/r2c40_bench.err
grep -o '"value": [0-9.]*' gpurun_out/r2c40_bench.json | head -2
grep breakdown gpurun_out/r2c40_bench.err
