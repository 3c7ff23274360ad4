cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_attention_gpu.py -q -m gpu -k "core_vs_oracle" 2>&1 | tail -40 > gpurun_out/r2c2_tests.log
tail -8 gpurun_out/r2c2_tests.log
timeout 120 python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 3 > gpurun_out/r2c2_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 0 -c 3 -f -o gpurun_out/r2c2_flash python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 1 > gpurun_out/r2c2_ncu.log 2>&1
cat gpurun_out/r2c2_plain.log; tail -3 gpurun_out/r2c2_ncu.log
