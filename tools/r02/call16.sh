#!/bin/bash
# r02 call 16: MMA issue loop — operand-kind branch hoisted out of the loop (default) vs additionally a single-thread loop with
# precomputed descriptors (F5_ISSUE1), against round 1 and the last commit before the FP8 mode, all on the same box
OUT=$PWD/gpurun_out
mkdir -p $OUT
lscpu | grep -E "Model name" | head -1
run() {  # dir tag flags [lib]
  (cd $1 && F5_LIB=$4 PYTHONPATH=. timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 2> $OUT/r02_c16_$2.err | tail -1 > $OUT/r02_c16_$2.json)
  python - $OUT/r02_c16_$2.json $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:28s} ms/step {d['ms_per_step']:.3f}  gemm {r.get('gemm_ms_per_step', 0):.2f} attn {r.get('attention', {}).get('ms_per_step', 0):.2f} other {r.get('other_ms_per_step', 0):.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
F5_LIB=$PWD/variants/libf5_issue1.so PYTHONPATH=. timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c16_kernels_issue1.log 2>&1; tail -2 $OUT/r02_c16_kernels_issue1.log
run variants/r01 r01 ""
run variants/t_884d77b c884d77b_fused "--no-configs"
run . hoist "--no-configs" $PWD/variants/libf5_hoist.so
run . issue1 "--no-configs" $PWD/variants/libf5_issue1.so
run . hoist_fp8 "--no-configs --fp8" $PWD/variants/libf5_hoist.so
run . issue1_fp8 "--no-configs --fp8" $PWD/variants/libf5_issue1.so
run . hoist_b64 "--no-configs --batch 64 --method midpoint --steps 2 --warmup 1" $PWD/variants/libf5_hoist.so
run . issue1_b64 "--no-configs --batch 64 --method midpoint --steps 2 --warmup 1" $PWD/variants/libf5_issue1.so
