cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_optim_gpu.py "tests/test_configs_gpu.py::test_config1_swin_tiny_default_windows_forward_backward" -q -m gpu -x --tb=short 2>&1 | tail -80 > gpurun_out/r2c5_tests.log
timeout 1200 python -m pytest "tests/test_configs_gpu.py::test_config1_swin_tiny_default_windows_forward_backward" -q -m gpu --tb=short 2>&1 | tail -40 >> gpurun_out/r2c5_tests.log
cat gpurun_out/r2c5_tests.log
