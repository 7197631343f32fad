#!/bin/bash
# same-box trajectory: each round-2 commit's own tree + bench (B=1, 10 steps), fused and unfused AdaLN
OUT=$PWD/gpurun_out
mkdir -p $OUT
lscpu | grep -E "Model name" | head -1
run() {  # dir tag flags
  (cd $1 && PYTHONPATH=. timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 2> $OUT/r02_c15_$2.err | tail -1 > $OUT/r02_c15_$2.json)
  python - $OUT/r02_c15_$2.json $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:28s} ms/step {d['ms_per_step']:.3f}  gemm {r.get('gemm_ms_per_step', 0):.2f} attn {r.get('attention', {}).get('ms_per_step', 0):.2f} other {r.get('other_ms_per_step', 0):.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
run variants/r01 r01 ""
for c in e38944c 96aee68 8a24954 884d77b; do
  run variants/t_$c ${c}_fused "--no-configs"
  run variants/t_$c ${c}_unfused "--no-configs --no-fused-adaln"
done
run . head_fused "--no-configs"
run . head_unfused "--no-configs --no-fused-adaln"
run variants/r01 r01_again ""
