#!/bin/bash
cd /root/repo
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2c41_dp2.json 2> gpurun_out/r2c41_dp2.err
grep -o '"value": [0-9.]*' gpurun_out/r2c41_dp2.json | head -3; tail -3 gpurun_out/r2c41_dp2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c4_swinL_kitti_infer --no-extras > gpurun_out/r2c41_c4_2.json 2> gpurun_out/r2c41_c4_2.err
grep -o '"value": [0-9.]*' gpurun_out/r2c41_c4_2.json | head -1
