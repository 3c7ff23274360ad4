set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r01_final_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_final_ref.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 2200 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 100 python tools/prof_attn_raw.py > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*ws -s 4 -c 2 -f -o gpurun_out/r01b_attn python tools/prof_attn_raw.py > gpurun_out/ncu_a.log 2>&1
timeout 100 python tools/prof_one_gemm.py gelu 2 > gpurun_out/plain_g.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r01b_gemm_gelu_st2 python tools/prof_one_gemm.py gelu 2 > gpurun_out/ncu_g.log 2>&1
timeout 100 python tools/prof_one_gemm.py plain 2 > gpurun_out/plain_g2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r01b_gemm_plain_st2 python tools/prof_one_gemm.py plain 2 > gpurun_out/ncu_g2.log 2>&1
timeout 100 python tools/prof_one_ln.py 2 > gpurun_out/plain_n.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_.*bf16 -s 2 -c 2 -f -o gpurun_out/r01b_ln_st2 python tools/prof_one_ln.py 2 > gpurun_out/ncu_n.log 2>&1
tail -2 gpurun_out/r01_final_bench.log | cut -c1-300
ls -la gpurun_out | tail -12
