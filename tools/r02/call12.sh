#!/bin/bash
# r02 call 12: FP8 mode on all four block GEMMs — tests, in-situ, bench A/B on the SAME box (bf16 vs fp8 level 1 vs level 2)
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c12_kernels.log 2>&1; rc=$?; tail -15 $OUT/r02_c12_kernels.log
if [ $rc -ne 0 ]; then echo "KERNEL TESTS FAILED"; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "fp8 or dit_forward or config1 or full_config2 or duration" > $OUT/r02_c12_parity.log 2>&1; rc=$?; tail -15 $OUT/r02_c12_parity.log
F5_FP8=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c12_insitu_fp8.log 2>&1; cat $OUT/r02_c12_insitu_fp8.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c12_b1_bf16.err | tail -1 > $OUT/r02_c12_b1_bf16.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --fp8 2> $OUT/r02_c12_b1_fp8.err | tail -1 > $OUT/r02_c12_b1_fp8.json
F5_FP8_LEVEL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --fp8 2> $OUT/r02_c12_b1_fp8l1.err | tail -1 > $OUT/r02_c12_b1_fp8l1.json
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 64 --method midpoint 2> $OUT/r02_c12_b64_bf16.err | tail -1 > $OUT/r02_c12_b64_bf16.json
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 64 --method midpoint --fp8 2> $OUT/r02_c12_b64_fp8.err | tail -1 > $OUT/r02_c12_b64_fp8.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_c12_b*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "gemm ms", round(r["gemm_ms_per_step"], 2), "attn ms", round(r["attention"]["ms_per_step"], 2), "other", round(r["other_ms_per_step"], 2), d["dtype"])
    except Exception as e:
        print(f, "ERR", e)
PY
