cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_optim_gpu.py tests/test_configs_gpu.py -q -m gpu --tb=short 2>&1 | tail -8
for wl in c2_ws24 c2_ws30 c1_swinT kitti_train void_train c4_swinL_kitti_infer c3_void_silog c5_micro; do
  echo "=== $wl"
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2c7_$wl.json 2> gpurun_out/r2c7_$wl.err || tail -5 gpurun_out/r2c7_$wl.err
  cut -c1-700 gpurun_out/r2c7_$wl.json
done
