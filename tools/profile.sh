#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full captures of the top kernels of one bench step.
# Outputs go to gpurun_out/ (summaries are copied into profiles/ by tools/summarize_profiles.py here).
set -x
OUT=gpurun_out
mkdir -p $OUT
# every launch of precompute + the first DiT evaluations (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $OUT/launches.csv \
    python bench.py --profile-run > $OUT/ncu_launches.log 2>&1
# top kernels, full set, 2 launches each, skipping the precompute launches
ncu --set full --clock-control none --import-source on -k regex:gemm -s 40 -c 6 -o $OUT/prof_gemm -f \
    python bench.py --profile-run > $OUT/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn -s 2 -c 2 -o $OUT/prof_attn -f \
    python bench.py --profile-run > $OUT/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_mod -s 4 -c 2 -o $OUT/prof_ln -f \
    python bench.py --profile-run > $OUT/ncu_ln.log 2>&1
ls -la $OUT
