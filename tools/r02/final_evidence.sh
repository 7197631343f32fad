#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, the bench lines (both arms), ncu launch list + full captures,
# compute-sanitizer memcheck / racecheck logs.  Summaries are copied into profiles/ by tools/summarize_profiles.py.
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,driver_version --format=csv > $OUT/final_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $OUT/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/final_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 2> $OUT/bench_final_b1.err | tail -1 > $OUT/bench_final_b1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_final_ref.err | tail -1 > $OUT/bench_final_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final_b1.json")); r = d["roofline"]
print("B1 ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(r["frac"], 3), "whole", round(r["whole_step"]["frac"], 3),
      "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2), "clocks", d["clocks"])
for k, v in d["configs"].items():
    rr = v["roofline"]; print("  ", k, round(v["ms_per_step"], 2), round(v["value"]), "gemm", round(rr["frac"], 3), "attn TF", round(rr["attention"]["achieved"]), "whole", round(rr["whole_step"]["frac"], 3))
print("ref", json.load(open("gpurun_out/bench_final_ref.json"))["value"])
PY
timeout 1200 bash tools/profile.sh > $OUT/profile.log 2>&1; ls $OUT/*.ncu-rep 2>/dev/null | wc -l
# memory / race checks of the tensor-core kernels (slow under the sanitizer: small subsets)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 500 -k "(gemm or attention) and not 40000 and not long" > $OUT/final_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/final_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 450 -k "test_gemm_qkv_rope_epilogue or test_gemm_gate_mask_residual_inplace or (fused_ln_producer and 700) or (test_attention and 937 and None)" > $OUT/final_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/final_racecheck.log
