cd $GRAFT_REPO_ROOT
timeout 120 python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 3 --bwd 0 > gpurun_out/r2c16_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 2 -c 1 -f -o gpurun_out/r2c16_flash_fwd python tools/prof_attn_raw.py --impl 1 --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 1 --bwd 0 > gpurun_out/r2c16_ncu.log 2>&1
cat gpurun_out/r2c16_plain.log
