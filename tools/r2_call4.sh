cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests/test_optim_gpu.py tests/test_configs_gpu.py tests/test_attention_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2c4_tests.log
tail -30 gpurun_out/r2c4_tests.log
