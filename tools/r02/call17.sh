#!/bin/bash
# r02 call 17: single-thread issue loops (GEMM default now; attention new) on one box, vs 884d77b
OUT=$PWD/gpurun_out
mkdir -p $OUT
lscpu | grep -E "Model name" | head -1
run() {  # dir tag flags [lib]
  (cd $1 && F5_LIB=$4 PYTHONPATH=. timeout 400 python bench.py --warmup 3 --no-cpu-baseline $3 2> $OUT/r02_c17_$2.err | tail -1 > $OUT/r02_c17_$2.json)
  python - $OUT/r02_c17_$2.json $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print(f"{sys.argv[2]:28s} ms/step {d['ms_per_step']:.3f}  gemm {r.get('gemm_ms_per_step', 0):.2f} attn {r.get('attention', {}).get('ms_per_step', 0):.2f} other {r.get('other_ms_per_step', 0):.2f}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
PYTHONPATH=. timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c17_kernels.log 2>&1; tail -2 $OUT/r02_c17_kernels.log
PYTHONPATH=. timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "dit_forward or config1 or full_config2 or full_config5 or ragged or fp8" > $OUT/r02_c17_parity.log 2>&1; tail -2 $OUT/r02_c17_parity.log
run variants/t_884d77b c884d77b "--no-configs --steps 10"
run . new "--no-configs --steps 10" $PWD/variants/libf5_new.so
run . attn0 "--no-configs --steps 10" $PWD/variants/libf5_attn0.so
run . new_fp8 "--no-configs --steps 10 --fp8" $PWD/variants/libf5_new.so
run . new_long "--no-configs --steps 3 --frames 5625 --ref-frames 499" $PWD/variants/libf5_new.so
run . attn0_long "--no-configs --steps 3 --frames 5625 --ref-frames 499" $PWD/variants/libf5_attn0.so
run . new_b64 "--no-configs --batch 64 --method midpoint --steps 2 --warmup 1" $PWD/variants/libf5_new.so
run . attn0_b64 "--no-configs --batch 64 --method midpoint --steps 2 --warmup 1" $PWD/variants/libf5_attn0.so
