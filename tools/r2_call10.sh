cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -15
timeout 600 python tools/prof_gemm.py --iters 3 > gpurun_out/r2c10_gemm.log 2>&1
tail -45 gpurun_out/r2c10_gemm.log
