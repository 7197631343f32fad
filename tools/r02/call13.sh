#!/bin/bash
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 -k "attention" > $OUT/r02_c13_attn.log 2>&1; tail -3 $OUT/r02_c13_attn.log
timeout 400 python tests/gpu_checks/check_generate_rtf.py 2>&1 | grep -E "GENERATE_RTF|Error|error" | tail -3 | tee $OUT/r02_c13_generate_rtf.log
