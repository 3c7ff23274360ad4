#!/bin/bash
cd /root/repo
for k in 1 2; do timeout 60 python -u tools/check_mma.py --bwd 1 --a 1 --b 4 2>&1 | grep -E "impl"; done
timeout 60 python -u tools/check_mma.py --bwd 1 --a 1 --b 4 --shift 0 2>&1 | grep -E "impl"
