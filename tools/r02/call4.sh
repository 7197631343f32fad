#!/bin/bash
# r02 call 4: TMA-store epilogue — kernel tests, parity, in-situ timelines, bench A/B (stages stop at the first failure)
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c4_kernels.log 2>&1; rc=$?; tail -12 $OUT/r02_c4_kernels.log
if [ $rc -ne 0 ]; then echo "KERNEL TESTS FAILED rc=$rc"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_pins.py -x -q -m gpu --timeout 300 > $OUT/r02_c4_parity.log 2>&1; rc=$?; tail -25 $OUT/r02_c4_parity.log
if [ $rc -ne 0 ]; then echo "PARITY TESTS FAILED rc=$rc"; exit 0; fi
F5_FUSED=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c4_insitu_fused.log 2>&1
F5_FUSED=0 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c4_insitu_unfused.log 2>&1
cat $OUT/r02_c4_insitu_fused.log $OUT/r02_c4_insitu_unfused.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/r02_c4_bench.err | tail -1 > $OUT/r02_c4_bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-fused-adaln 2> $OUT/r02_c4_bench_unfused.err | tail -1 > $OUT/r02_c4_bench_unfused.json
tail -c 400 $OUT/r02_c4_bench.err; tail -c 300 $OUT/r02_c4_bench_unfused.err
python - <<'PY'
import json
for f in ("r02_c4_bench", "r02_c4_bench_unfused"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        r = d["roofline"]
        print(f, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 3),
              "gemm ms", round(r["gemm_ms_per_step"], 2), "attn ms", round(r["attention"]["ms_per_step"], 2), "other", round(r["other_ms_per_step"], 2), "launches", d["launches_per_step"])
        for k, v in d.get("configs", {}).items():
            rr = v["roofline"]
            print("  ", k, "ms/step", round(v["ms_per_step"], 2), "value", round(v["value"]), "gemm frac", round(rr["frac"], 3), "attn TF", round(rr["attention"]["achieved"]), "whole", round(rr["whole_step"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
