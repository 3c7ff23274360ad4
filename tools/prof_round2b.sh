# Round-2 final evidence: GPU tests, default bench line (with extras), every other workload, launch list of a graph
# replay, ncu --set full of the warp-MMA attention kernels.
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -5 > gpurun_out/r02f_tests.log; tail -3 gpurun_out/r02f_tests.log
timeout 900 python bench.py --breakdown > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; grep breakdown gpurun_out/r02f_bench.err; grep -o '"value": [0-9.]*' gpurun_out/r02f_bench.json | head -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_reference.json 2> gpurun_out/r02f_reference.err; head -c 400 gpurun_out/r02f_reference.json
for w in c2_ws24 c2_ws30 c1_swinT kitti_train void_train c4_swinL_kitti_infer c3_void_silog; do timeout 600 python bench.py --workload $w --no-extras > gpurun_out/r02f_$w.json 2> gpurun_out/r02f_$w.err; grep -o '"value": [0-9.]*' gpurun_out/r02f_$w.json | head -1; done
timeout 900 python bench.py --workload c5_micro > gpurun_out/r02f_c5_micro.json 2> gpurun_out/r02f_c5_micro.err; tail -c 300 gpurun_out/r02f_c5_micro.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/plain_l.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 2400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_l.log 2>&1
timeout 100 python tools/check_mma.py --iters 1 --bwd 1 --a 1 --b 1 > gpurun_out/plain_m.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma -s 2 -c 2 -f -o gpurun_out/r02f_attn_mma_ws12 python tools/check_mma.py --iters 1 --bwd 1 --a 1 --b 1 > gpurun_out/ncu_m.log 2>&1
tail -2 gpurun_out/plain_m.log
