#!/bin/bash
# r02 call 1: GPU tests at HEAD, baseline of the r01 kernels on configs 5 and 3, GEMM variant sweep vs cuBLAS, in-situ timelines
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $OUT/r02_c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_c1_pytest.log 2>&1
tail -5 $OUT/r02_c1_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --frames 5625 --ref-frames 499 --no-cpu-baseline 2> $OUT/r02_c1_long.err | tail -1 > $OUT/r02_c1_long.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 64 --method midpoint --no-cpu-baseline 2> $OUT/r02_c1_b64.err | tail -1 > $OUT/r02_c1_b64.json
timeout 300 python tests/gpu_checks/check_insitu.py > $OUT/r02_c1_insitu.log 2>&1
timeout 300 python tests/gpu_checks/check_gemm2.py > $OUT/r02_c1_gemm2.log 2>&1
tail -c 700 $OUT/r02_c1_long.json; echo; tail -c 300 $OUT/r02_c1_long.err; tail -c 700 $OUT/r02_c1_b64.json; echo; tail -30 $OUT/r02_c1_insitu.log; grep time $OUT/r02_c1_gemm2.log
