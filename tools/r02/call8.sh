#!/bin/bash
# r02 call 8: row-sticky QKV tile walk + read-only final store wait — tests, in-situ, bench B=1 / B=64
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c8_kernels.log 2>&1; rc=$?; tail -6 $OUT/r02_c8_kernels.log
if [ $rc -ne 0 ]; then echo "KERNEL TESTS FAILED rc=$rc"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "dit_forward or config1 or fused_adaln or full_config2 or ragged or bucketing or golden_fixture" > $OUT/r02_c8_parity.log 2>&1; rc=$?; tail -6 $OUT/r02_c8_parity.log
if [ $rc -ne 0 ]; then echo "PARITY TESTS FAILED rc=$rc"; exit 0; fi
F5_FUSED=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c8_insitu_fused.log 2>&1; cat $OUT/r02_c8_insitu_fused.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c8_b1.err | tail -1 > $OUT/r02_c8_b1.json
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 64 --method midpoint 2> $OUT/r02_c8_b64.err | tail -1 > $OUT/r02_c8_b64.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_c8_b*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "gemm frac", round(r["frac"], 3),
              "gemm ms", round(r["gemm_ms_per_step"], 2), "attn ms", round(r["attention"]["ms_per_step"], 2), "other", round(r["other_ms_per_step"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 $OUT/r02_c8_b1.err
