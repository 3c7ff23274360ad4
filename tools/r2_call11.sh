cd $GRAFT_REPO_ROOT
timeout 100 python tools/prof_one_gemm.py plain 2 > gpurun_out/r2c11_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/r2c11_gemm_plain_st2 python tools/prof_one_gemm.py plain 2 > gpurun_out/r2c11_ncu.log 2>&1
cat gpurun_out/r2c11_plain.log | tail -3
