#!/bin/bash
# r02 call 22: variant sweep again (F5_TUNE) — the issue-loop fix changed every variant's main loop by 20-35 %
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
i=0
for t in "" "qkv=2:256" "ff1=2:256" "qkv=2:256,ff1=2:256" "out=1:64,ff2=1:64" "ff2=1:64" "qkv=1:128" ""; do
  i=$((i+1))
  unset F5_TUNE
  if [ -n "$t" ]; then export F5_TUNE="$t"; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c22_$i.err | tail -1 > $OUT/r02_c22_$i.json
  python - "$t" $OUT/r02_c22_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d["roofline"]
    print(f"{sys.argv[1] or 'default':22s} ms/step {d['ms_per_step']:.3f} gemm ms {r['gemm_ms_per_step']:.2f} attn {r['attention']['ms_per_step']:.2f} other {r['other_ms_per_step']:.2f}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
