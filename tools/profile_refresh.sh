#!/bin/bash
# Reduced refresh of the ncu evidence after a kernel change (tools/profile.sh is the full set): launch list, one cold
# and one warm --set full capture of the block's four GEMMs (consecutive launches of one DiT block), one of the attention.
# Mangled names: gemm2 <BN, stages, act, out_bf16, rope, fp8> / gemm <BN, stages, act, out_bf16, rope, fp8, resid>.
set -x
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/prof_*.ncu-rep $OUT/launches.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --profile-run > $OUT/ncu_launches.log 2>&1
GEMMS="regex:gemm2_bf16_tn_kernelILi128ELi6ELi0ELb1ELb1|gemm_bf16_tn_kernelILi128ELi6ELi0ELb0ELb0|gemm_bf16_tn_kernelILi128ELi3ELi1ELb1ELb0"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "$GEMMS" -s 40 -c 4 -o $OUT/prof_gemm_block -f \
    python bench.py --profile-run > $OUT/ncu_gemm_block.log 2>&1
ncu --set full --clock-control none --cache-control none --kernel-name-base mangled -k "$GEMMS" -s 40 -c 4 -o $OUT/prof_gemm_block_warm -f \
    python bench.py --profile-run > $OUT/ncu_gemm_warm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn -s 30 -c 1 -o $OUT/prof_attn -f \
    python bench.py --profile-run > $OUT/ncu_attn.log 2>&1
ls -la $OUT | grep -E "prof_|launches"
