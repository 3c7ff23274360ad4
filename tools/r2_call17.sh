cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2c17_dp2.json 2> gpurun_out/r2c17_dp2.err
tail -3 gpurun_out/r2c17_dp2.err | cut -c1-300; cut -c1-600 gpurun_out/r2c17_dp2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --workload c4_swinL_kitti_infer > gpurun_out/r2c17_c4_2.json 2> gpurun_out/r2c17_c4_2.err
tail -2 gpurun_out/r2c17_c4_2.err | cut -c1-300; cut -c1-400 gpurun_out/r2c17_c4_2.json
