#!/bin/bash
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
F5_FUSED=1 timeout 300 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c3_insitu_fused.log 2>&1
F5_FUSED=0 timeout 300 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c3_insitu_unfused.log 2>&1
cat $OUT/r02_c3_insitu_fused.log $OUT/r02_c3_insitu_unfused.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bucketing or generate_end_to_end" > $OUT/r02_c3_tests.log 2>&1; tail -15 $OUT/r02_c3_tests.log
