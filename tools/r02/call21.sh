#!/bin/bash
# r02 call 21: in-situ GEMM phases after the elect.sync change (bf16 and FP8 mode)
OUT=$PWD/gpurun_out
mkdir -p $OUT
PYTHONPATH=. timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c21_insitu_bf16.log 2>&1; cat $OUT/r02_c21_insitu_bf16.log | tail -8
F5_FP8=1 PYTHONPATH=. timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c21_insitu_fp8.log 2>&1; cat $OUT/r02_c21_insitu_fp8.log | tail -8
