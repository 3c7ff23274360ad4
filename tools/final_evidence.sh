#!/bin/bash
# Final evidence of a round on one B200: GPU test log + one bench line per workload under gpurun_out/final/.
#   gpurun --timeout 2400 -- 'bash tools/final_evidence.sh'
out=gpurun_out/final
mkdir -p $out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/gpu_tests.log
tail -3 $out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
python bench.py --steps 20 --warmup 3 --breakdown > $out/c2_ws12_n1.json 2> $out/c2_ws12_n1_breakdown.txt
python bench.py --impl reference --steps 2 --warmup 1 > $out/c2_ws12_reference_arm.json 2> $out/reference_arm.err
for w in c2_ws24 c2_ws30 c1_swinT kitti_train void_train; do
  python bench.py --workload $w --steps 8 --warmup 3 --no-extras --breakdown > $out/$w.json 2> $out/$w.err
done
for w in c4_swinL_kitti_infer c3_void_silog c3_void_encoder c5_micro; do
  python bench.py --workload $w --steps 10 --warmup 3 > $out/$w.json 2> $out/$w.err
done
python bench.py --workload c3_void_encoder --dtype fp32 --steps 5 --warmup 3 > $out/c3_void_encoder_fp32.json 2> $out/c3_void_encoder_fp32.err
python tools/prof_mha.py 16 > $out/mha_timing.log 2>&1
for f in $out/*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d.get('value', 0), 1), d.get('unit'), 'ms/step', round(d.get('ms_per_step', 0), 2), 'e2e', round((d.get('e2e') or {}).get('value', 0), 1))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
