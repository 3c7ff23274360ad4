# Round-2 evidence: full GPU tests, bench, launch list of a graph replay, ncu --set full of the dominant kernels.
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -5 > gpurun_out/r02_tests.log; tail -3 gpurun_out/r02_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --breakdown --no-extras > gpurun_out/r02_bench_noextras.json 2> gpurun_out/r02_bench_noextras.err; grep breakdown gpurun_out/r02_bench_noextras.err
timeout 300 python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/plain_l.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 2400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_l.log 2>&1
timeout 100 python tools/prof_attn_raw.py --bwd 0 > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 4 -c 1 -f -o gpurun_out/r02_attn_fwd_ws12 python tools/prof_attn_raw.py --bwd 0 > gpurun_out/ncu_a.log 2>&1
timeout 100 python tools/prof_attn_raw.py > gpurun_out/plain_a2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_ws -s 4 -c 1 -f -o gpurun_out/r02_attn_bwd_ws12 python tools/prof_attn_raw.py > gpurun_out/ncu_a2.log 2>&1
timeout 100 python tools/prof_attn_raw.py --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 2 > gpurun_out/plain_a3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 9 -c 3 -f -o gpurun_out/r02_attn_flash_ws24 python tools/prof_attn_raw.py --B 8 --H 120 --C 128 --ws 24 --shift 12 --iters 2 > gpurun_out/ncu_a3.log 2>&1
timeout 100 python tools/prof_one_gemm.py gelu 2 > gpurun_out/plain_g.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r02_gemm_gelu_st2 python tools/prof_one_gemm.py gelu 2 > gpurun_out/ncu_g.log 2>&1
timeout 100 python tools/prof_one_gemm.py plain 2 > gpurun_out/plain_g2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r02_gemm_plain_st2 python tools/prof_one_gemm.py plain 2 > gpurun_out/ncu_g2.log 2>&1
timeout 100 python tools/prof_one_ln.py 2 > gpurun_out/plain_n.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_.*bf16 -s 2 -c 2 -f -o gpurun_out/r02_ln_st2 python tools/prof_one_ln.py 2 > gpurun_out/ncu_n.log 2>&1
cat gpurun_out/plain_a2.log | tail -1
