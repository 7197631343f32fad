#!/bin/bash
# r02 call 6: in-situ sweep of kernel variants for the block's GEMMs (F5_TUNE), B=1
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
i=0
for t in "" "qkv=1:128" "qkv=2:256" "qkv=2:128" "ff1=2:256" "ff1=2:128" "ff1=1:64" "out=1:64,ff2=1:64" "ff2=1:64" "PREFETCH0" "PDL0"; do
  i=$((i+1))
  unset F5_TUNE F5_PREFETCH F5_PDL
  if [ "$t" = "PREFETCH0" ]; then export F5_PREFETCH=0; elif [ "$t" = "PDL0" ]; then export F5_PDL=0; elif [ -n "$t" ]; then export F5_TUNE="$t"; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c6_$i.err | tail -1 > $OUT/r02_c6_$i.json
  python - "$t" $OUT/r02_c6_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d["roofline"]
    print(f"{sys.argv[1] or 'default':22s} ms/step {d['ms_per_step']:.3f} gemm ms {r['gemm_ms_per_step']:.2f} attn {r['attention']['ms_per_step']:.2f} other {r['other_ms_per_step']:.2f}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
