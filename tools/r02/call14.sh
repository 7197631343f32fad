#!/bin/bash
# same-box comparison: round-1 kernels (tree at commit 0dd48cd under variants/r01) vs HEAD, interleaved
OUT=$PWD/gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,serial,uuid,pci.bus_id,vbios_version --format=csv > $OUT/r02_c14_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/r02_c14_gpu.txt
for rep in 1 2; do
  (cd variants/r01 && PYTHONPATH=. timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/r02_c14_r01_$rep.err | tail -1 > $OUT/r02_c14_r01_$rep.json)
  PYTHONPATH=. timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2> $OUT/r02_c14_head_$rep.err | tail -1 > $OUT/r02_c14_head_$rep.json
done
PYTHONPATH=. timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-fused-adaln 2> $OUT/r02_c14_head_unfused.err | tail -1 > $OUT/r02_c14_head_unfused.json
cat $OUT/r02_c14_gpu.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_c14_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
