#!/bin/bash
# localise the remaining GEMM regression vs 884d77b: in-situ phases of both on one box
OUT=$PWD/gpurun_out
mkdir -p $OUT
lscpu | grep -E "Model name" | head -1
cp tests/gpu_checks/check_insitu2.py variants/t_884d77b/tests/gpu_checks/check_insitu2_new.py
(cd variants/t_884d77b && PYTHONPATH=. F5_FUSED=1 timeout 200 python tests/gpu_checks/check_insitu2.py) > $OUT/r02_c18_insitu_884.log 2>&1; cat $OUT/r02_c18_insitu_884.log
F5_LIB=$PWD/variants/libf5_attn0.so PYTHONPATH=. F5_FUSED=1 timeout 200 python tests/gpu_checks/check_insitu2.py > $OUT/r02_c18_insitu_head.log 2>&1; cat $OUT/r02_c18_insitu_head.log
