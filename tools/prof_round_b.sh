set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r01_final_bench2.log 2>&1; tail -1 gpurun_out/r01_final_bench2.log | cut -c1-260
timeout 100 python tools/prof_attn_raw.py > gpurun_out/plain_a2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_ws -s 4 -c 1 -f -o gpurun_out/r01b_attn_bwd python tools/prof_attn_raw.py > gpurun_out/ncu_a2.log 2>&1
tail -2 gpurun_out/plain_a2.log
