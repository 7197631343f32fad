#!/bin/bash
# r02 call 24: where a chunk of the one-wave epilogue spends its time (five probe builds of gemm.cu)
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/r02_c24_epi_probe.log
for n in 1 2 3 4 5; do
  F5_LIB=$PWD/variants/probe/libf5b200_p$n.so F5_PROBE_LEVEL=$n PYTHONPATH=. timeout 120 python tests/gpu_checks/check_epi_probe.py >> $OUT/r02_c24_epi_probe.log 2>&1
done
cat $OUT/r02_c24_epi_probe.log | tail -20
