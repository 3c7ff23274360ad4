#!/bin/bash
# HMMA / ldmatrix / movmatrix rates on B200 + LayerNorm stream tests after the materialize_grads change + a short bench
cd /root/repo
timeout 120 ./build/hmma_rate > gpurun_out/r2c21_hmma.log 2>&1
cat gpurun_out/r2c21_hmma.log
timeout 600 python -m pytest tests/test_layernorm_gpu.py tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r2c21_tests.log
cat gpurun_out/r2c21_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err
grep -o '"value": [0-9.]*' gpurun_out/r2c21_bench.json | head -2
grep breakdown gpurun_out/r2c21_bench.err
