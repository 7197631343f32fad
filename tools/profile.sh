#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full captures of the top kernels of one bench step.
# Outputs go to gpurun_out/ (summaries are copied into profiles/ by tools/summarize_profiles.py here).
set -x
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/prof_*.ncu-rep $OUT/launches.csv
# every launch of precompute + the first DiT evaluations (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $OUT/launches.csv \
    python bench.py --profile-run > $OUT/ncu_launches.log 2>&1
# top kernels, full set, skipping the precompute launches: one block's QKV (pair kernel), out-proj, FF1, FF2
ncu --set full --clock-control none --import-source on -k regex:gemm -s 44 -c 5 -o $OUT/prof_gemm -f \
    python bench.py --profile-run > $OUT/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn -s 2 -c 2 -o $OUT/prof_attn -f \
    python bench.py --profile-run > $OUT/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_mod -s 4 -c 2 -o $OUT/prof_ln -f \
    python bench.py --profile-run > $OUT/ncu_ln.log 2>&1
ls -la $OUT
