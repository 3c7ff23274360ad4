# End-of-round evidence: launch list of a graph replay + ncu --set full of the dominant kernels (run under gpurun).
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 2200 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 100 python tools/prof_attn_raw.py --bwd 0 > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 4 -c 1 -f -o gpurun_out/r01b_attn python tools/prof_attn_raw.py --bwd 0 > gpurun_out/ncu_a.log 2>&1
timeout 100 python tools/prof_attn_raw.py > gpurun_out/plain_a2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_ws -s 4 -c 1 -f -o gpurun_out/r01b_attn_bwd python tools/prof_attn_raw.py > gpurun_out/ncu_a2.log 2>&1
timeout 100 python tools/prof_one_gemm.py gelu 2 > gpurun_out/plain_g.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r01b_gemm_gelu_st2 python tools/prof_one_gemm.py gelu 2 > gpurun_out/ncu_g.log 2>&1
timeout 100 python tools/prof_one_gemm.py plain 2 > gpurun_out/plain_g2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r01b_gemm_plain_st2 python tools/prof_one_gemm.py plain 2 > gpurun_out/ncu_g2.log 2>&1
timeout 100 python tools/prof_one_ln.py 2 > gpurun_out/plain_n.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_.*bf16 -s 2 -c 2 -f -o gpurun_out/r01b_ln_st2 python tools/prof_one_ln.py 2 > gpurun_out/ncu_n.log 2>&1
cat gpurun_out/plain_a2.log | tail -1
