#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 > gpurun_out/r2c32_tests.log
cat gpurun_out/r2c32_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err
grep -o '"value": [0-9.]*' gpurun_out/r2c32_bench.json | head -2
grep breakdown gpurun_out/r2c32_bench.err
