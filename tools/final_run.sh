#!/bin/bash
# Round-end evidence on the GPU box: ncu launch list + full captures (tools/profile.sh) and the bench lines.
export PYTHONPATH=.
OUT=gpurun_out
timeout 1200 bash tools/profile.sh > $OUT/profile.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 2> $OUT/bench_b1.err | tail -1 > $OUT/bench_final_b1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_ref.err | tail -1 > $OUT/bench_final_ref.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 64 --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/bench_final_b64_euler.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 64 --method midpoint --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/bench_final_b64_midpoint.json
for f in b1 ref b64_euler b64_midpoint; do python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_final_$f.json")); print("$f", d.get("ms_per_step"), d.get("value"), d.get("e2e",{}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
