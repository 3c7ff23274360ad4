cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/r2c6_tests.log
tail -25 gpurun_out/r2c6_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 --breakdown > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
tail -5 gpurun_out/r2c6_bench.err; cut -c1-3000 gpurun_out/r2c6_bench.json
