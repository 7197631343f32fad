#!/bin/bash
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python tests/gpu_checks/check_fp8_gemm.py > $OUT/r02_c11_fp8_gemm.log 2>&1; cat $OUT/r02_c11_fp8_gemm.log
F5_FP8=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c11_insitu_fp8.log 2>&1; cat $OUT/r02_c11_insitu_fp8.log
F5_FP8=0 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c11_insitu_bf16.log 2>&1; cat $OUT/r02_c11_insitu_bf16.log
