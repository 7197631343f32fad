#!/bin/bash
# r02 call 9: second TMA producer thread (A / B tiles requested by two threads) vs single producer
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c9_kernels.log 2>&1; rc=$?; tail -6 $OUT/r02_c9_kernels.log
if [ $rc -ne 0 ]; then echo "KERNEL TESTS FAILED rc=$rc"; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "dit_forward or config1 or full_config2 or ragged" > $OUT/r02_c9_parity.log 2>&1; rc=$?; tail -6 $OUT/r02_c9_parity.log
if [ $rc -ne 0 ]; then echo "PARITY TESTS FAILED rc=$rc"; exit 0; fi
F5_FUSED=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c9_insitu.log 2>&1; cat $OUT/r02_c9_insitu.log
for v in new cur; do
  if [ $v = cur ]; then export F5_LIB=$PWD/variants/libf5_cur.so; else unset F5_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c9_b1_$v.err | tail -1 > $OUT/r02_c9_b1_$v.json
  timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 64 --method midpoint 2> $OUT/r02_c9_b64_$v.err | tail -1 > $OUT/r02_c9_b64_$v.json
done
unset F5_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_c9_b*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "gemm frac", round(r["frac"], 3),
              "gemm ms", round(r["gemm_ms_per_step"], 2), "attn ms", round(r["attention"]["ms_per_step"], 2), "other", round(r["other_ms_per_step"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
