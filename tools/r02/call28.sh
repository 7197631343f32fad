#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 150 python tests/gpu_checks/check_gemm_vs_cublas.py > $OUT/r02_c28_gemm_vs_cublas.log 2>&1; tail -10 $OUT/r02_c28_gemm_vs_cublas.log
