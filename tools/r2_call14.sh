cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_attention_gpu.py tests/test_configs_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -6
{
timeout 120 python tools/prof_attn_raw.py --impl 2 --B 48 --H 30 --C 512 --ws 12 --bwd 0
timeout 120 python tools/prof_attn_raw.py --impl 2 --B 48 --H 120 --C 128 --ws 12 --shift 6 --bwd 0
timeout 120 python tools/prof_attn_raw.py --impl 2 --B 48 --H 60 --C 256 --ws 8 --shift 4 --bwd 0
timeout 300 python tools/prof_attn_raw.py --impl 1 --B 48 --H 120 --C 128 --ws 24 --shift 12 --iters 3 --bwd 0
timeout 300 python tools/prof_attn_raw.py --impl 1 --B 48 --H 120 --C 128 --ws 30 --shift 15 --iters 3 --bwd 0
timeout 300 python tools/prof_attn_raw.py --impl 1 --B 48 --H 60 --C 256 --ws 16 --shift 8 --iters 3 --bwd 0
} > gpurun_out/r2c14_timing.log 2>&1
cat gpurun_out/r2c14_timing.log
