#!/bin/bash
# final check at HEAD: smoke(), full GPU suite, the bench line of both arms
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $OUT/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/final_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 2> $OUT/bench_final_b1.err | tail -1 > $OUT/bench_final_b1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_final_ref.err | tail -1 > $OUT/bench_final_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final_b1.json")); r = d["roofline"]
print("B1 ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 3), "whole", round(r["whole_step"]["frac"], 3),
      "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2), "clocks", d["clocks"], "traffic", r.get("traffic"), r.get("traffic_warm"))
for k, v in d["configs"].items():
    rr = v["roofline"]; print("  ", k, round(v["ms_per_step"], 2), round(v["value"]), "gemm", round(rr["frac"], 3), "attn TF", round(rr["attention"]["achieved"]), "whole", round(rr["whole_step"]["frac"], 3))
print("ref", json.load(open("gpurun_out/bench_final_ref.json"))["value"])
PY
tail -c 300 $OUT/bench_final_b1.err
