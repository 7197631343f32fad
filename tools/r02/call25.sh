#!/bin/bash
# r02 call 25: epilogue arithmetic (LDS/STS instead of generic LD/ST, no mean term for non-consumers, 4-way partial
# sums) and a 7-stage ring for the one-wave kernel (variants/st7) — anchor: commit 5d8fac5 on the same box
OUT=$PWD/gpurun_out
mkdir -p $OUT
run() {  # dir tag flags [lib]
  (cd $1 && F5_LIB=$4 PYTHONPATH=. timeout 400 python bench.py --warmup 3 --no-cpu-baseline $3 2> $OUT/r02_c25_$2.err | tail -1 > $OUT/r02_c25_$2.json)
  python - $OUT/r02_c25_$2.json $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:28s} ms/step {d['ms_per_step']:.3f}  gemm {r.get('gemm_ms_per_step', 0):.2f} attn {r.get('attention', {}).get('ms_per_step', 0):.2f} other {r.get('other_ms_per_step', 0):.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
PYTHONPATH=. timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c25_kernels.log 2>&1; tail -2 $OUT/r02_c25_kernels.log
grep -q "passed" $OUT/r02_c25_kernels.log && ! grep -q "failed\|error" $OUT/r02_c25_kernels.log || { echo "kernel tests not green: stop"; tail -30 $OUT/r02_c25_kernels.log; exit 1; }
F5_LIB=$PWD/variants/st7/libf5b200.so PYTHONPATH=. timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 -k gemm > $OUT/r02_c25_kernels_st7.log 2>&1; tail -2 $OUT/r02_c25_kernels_st7.log
run variants/t_5d8fac5 c5d8fac5 "--no-configs --steps 10" ""
run . head "--no-configs --steps 10" ""
if grep -q "passed" $OUT/r02_c25_kernels_st7.log && ! grep -q "failed\|error" $OUT/r02_c25_kernels_st7.log; then
  run . head_st7 "--no-configs --steps 10" $PWD/variants/st7/libf5b200.so
  run . head_st7_fp8 "--no-configs --steps 10 --fp8" $PWD/variants/st7/libf5b200.so
fi
run . head_fp8 "--no-configs --steps 10 --fp8" ""
run . head_again "--no-configs --steps 10" ""
PYTHONPATH=. timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "dit_forward or config1 or full_config2 or ragged or fp8 or bucketing or fused" > $OUT/r02_c25_parity.log 2>&1; tail -2 $OUT/r02_c25_parity.log
