cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_layernorm_gpu.py -q -m gpu --tb=short 2>&1 | tail -8
