#!/bin/bash
# r02 call 27: new QKV tile rule (128-wide pair tiles), prologue stamps off the producer thread, row-mask branch; ff1=2:128 override; B=64
OUT=$PWD/gpurun_out
mkdir -p $OUT
run() {  # dir tag flags [lib]
  (cd $1 && F5_LIB=$4 PYTHONPATH=. timeout 400 python bench.py --warmup 3 --no-cpu-baseline $3 2> $OUT/r02_c27_$2.err | tail -1 > $OUT/r02_c27_$2.json)
  python - $OUT/r02_c27_$2.json $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:28s} ms/step {d['ms_per_step']:.3f}  gemm {r.get('gemm_ms_per_step', 0):.2f} attn {r.get('attention', {}).get('ms_per_step', 0):.2f} other {r.get('other_ms_per_step', 0):.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
PYTHONPATH=. timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c27_kernels.log 2>&1; tail -2 $OUT/r02_c27_kernels.log
grep -q "passed" $OUT/r02_c27_kernels.log && ! grep -q "failed\|error" $OUT/r02_c27_kernels.log || { echo "kernel tests not green: stop"; tail -30 $OUT/r02_c27_kernels.log; exit 1; }
run variants/t_5d8fac5 c5d8fac5 "--no-configs --steps 10" ""
run . head "--no-configs --steps 10" ""
run . head_fp8 "--no-configs --steps 10 --fp8" ""
(export F5_TUNE="ff1=2:128"; run . head_ff1_2_128 "--no-configs --steps 10" "")
run . head_b64 "--no-configs --steps 2 --batch 64 --method midpoint" ""
run . head_again "--no-configs --steps 10" ""
PYTHONPATH=. timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "dit_forward or config1 or full_config2 or ragged or fp8 or bucketing or fused" > $OUT/r02_c27_parity.log 2>&1; tail -2 $OUT/r02_c27_parity.log
