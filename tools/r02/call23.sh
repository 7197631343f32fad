#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 200 python tests/gpu_checks/check_epi_scaling.py > $OUT/r02_c23_epi_scaling.log 2>&1; cat $OUT/r02_c23_epi_scaling.log | tail -14
