#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full captures of the top kernels of one bench step.
# Outputs go to gpurun_out/ (summaries are copied into profiles/ by tools/summarize_profiles.py here).
# `bench.py --profile-run` = mel front-end + 2-point sample + Vocos, then ONE eager step of the bench workload.
set -x
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/prof_*.ncu-rep $OUT/launches.csv
# every launch of the e2e prologue + precompute + the first DiT evaluations (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --profile-run > $OUT/ncu_launches.log 2>&1
# tensor-core kernels of one DiT block, full set: QKV (pair kernel, RoPE + fused-LN consumer), out-proj / FF2 (one-wave
# kernel, fused-LN producer), FF1 (GELU, fused-LN consumer), attention
# (template arguments are only part of the MANGLED name: <BN, stages, act, out_bf16, rope> = ILi..ELi..ELi..ELb..ELb..E)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm2_bf16_tn_kernelILi192 -s 30 -c 2 -o $OUT/prof_gemm_qkv -f \
    python bench.py --profile-run > $OUT/ncu_gemm_qkv.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_bf16_tn_kernelILi128ELi6ELi0ELb0ELb0 -s 60 -c 2 -o $OUT/prof_gemm_out_ff2 -f \
    python bench.py --profile-run > $OUT/ncu_gemm_out.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_bf16_tn_kernelILi128ELi3ELi1ELb1ELb0 -s 30 -c 1 -o $OUT/prof_gemm_ff1 -f \
    python bench.py --profile-run > $OUT/ncu_gemm_ff1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn -s 30 -c 2 -o $OUT/prof_attn -f \
    python bench.py --profile-run > $OUT/ncu_attn.log 2>&1
# the same GEMMs without ncu's cache flush between replays (what the kernel sees inside a step: activations L2-resident)
ncu --set full --clock-control none --cache-control none --kernel-name-base mangled -k "regex:gemm_bf16_tn_kernelILi128ELi6ELi0ELb0ELb0|gemm2_bf16_tn_kernelILi192" -s 90 -c 3 -o $OUT/prof_gemm_warm -f \
    python bench.py --profile-run > $OUT/ncu_gemm_warm.log 2>&1
# HBM / FFT kernels: achieved DRAM GB/s
ncu --set full --clock-control none -k "regex:mel_kernel|istft|dwconv7|grn_|cfg_ode_update|text_embed|cast_pad|concat_cond|time_mlp|ln_tab_prep|ln_mod" -c 40 -o $OUT/prof_hbm -f \
    python bench.py --profile-run > $OUT/ncu_hbm.log 2>&1
ls -la $OUT
