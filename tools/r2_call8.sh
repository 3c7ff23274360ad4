cd $GRAFT_REPO_ROOT
timeout 120 python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 3 > gpurun_out/r2c8_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 3 -c 3 -f -o gpurun_out/r2c8_flash_bwd python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 1 > gpurun_out/r2c8_ncu.log 2>&1
cat gpurun_out/r2c8_plain.log; tail -2 gpurun_out/r2c8_ncu.log
