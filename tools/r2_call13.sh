cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -8
timeout 900 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err
grep breakdown gpurun_out/r2c13_bench.err; cut -c1-400 gpurun_out/r2c13_bench.json
