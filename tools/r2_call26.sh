#!/bin/bash
cd /root/repo
L=gpurun_out/r2c26.log
: > $L
run() { echo "== $*" >> $L; timeout 60 python -u tools/check_mma.py "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run --B 2 --H 24 --C 64 --shift 0 --bwd 1 --iters 2
run --B 64 --H 24 --C 64 --shift 0 --bwd 1 --iters 2
run --B 2 --H 30 --C 64 --shift 0 --bwd 1 --iters 2
run --B 2 --H 30 --C 64 --shift 0 --bwd 0 --iters 2
echo "== sanitizer" >> $L
timeout 240 compute-sanitizer --tool memcheck python -u tools/check_mma.py --B 2 --H 30 --C 64 --shift 0 --bwd 1 --iters 1 --b 4 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         " | head -40 >> $L
cat $L
