#!/bin/bash
# r02 call 2: fused AdaLN (LN by linearity) — kernel tests, parity incl. the full-size goldens, bench A/B
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > $OUT/r02_c2_kernels.log 2>&1; tail -15 $OUT/r02_c2_kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_ref_pins.py -q -m gpu > $OUT/r02_c2_parity.log 2>&1; tail -25 $OUT/r02_c2_parity.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/r02_c2_bench.err | tail -1 > $OUT/r02_c2_bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-fused-adaln 2> $OUT/r02_c2_bench_unfused.err | tail -1 > $OUT/r02_c2_bench_unfused.json
tail -c 400 $OUT/r02_c2_bench.err; tail -c 300 $OUT/r02_c2_bench_unfused.err
python - <<'PY'
import json
for f in ("r02_c2_bench", "r02_c2_bench_unfused"):
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
