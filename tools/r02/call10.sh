#!/bin/bash
# r02 call 10: FP8 mode tests + bench, racecheck after the fix, ncu GEMM captures
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 90 > $OUT/r02_c10_kernels.log 2>&1; rc=$?; tail -15 $OUT/r02_c10_kernels.log
if [ $rc -eq 0 ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "fp8 or dit_forward or config1 or full_config2" > $OUT/r02_c10_parity.log 2>&1; rc=$?; tail -15 $OUT/r02_c10_parity.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --fp8 2> $OUT/r02_c10_b1_fp8.err | tail -1 > $OUT/r02_c10_b1_fp8.json
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 64 --method midpoint --fp8 2> $OUT/r02_c10_b64_fp8.err | tail -1 > $OUT/r02_c10_b64_fp8.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_c10_b*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "gemm ms", round(r["gemm_ms_per_step"], 2), "attn ms", round(r["attention"]["ms_per_step"], 2), "other", round(r["other_ms_per_step"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 400 $OUT/r02_c10_b1_fp8.err
fi
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 450 -k "test_gemm_qkv_rope_epilogue or test_gemm_gate_mask_residual_inplace or (fused_ln_producer and 700) or (test_attention and 937 and None)" > $OUT/r02_c10_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" $OUT/r02_c10_racecheck.log; tail -3 $OUT/r02_c10_racecheck.log
# GEMM captures (kernel names with template arguments need the mangled base)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm2_bf16_tn_kernelILi192 -s 30 -c 2 -o $OUT/prof_gemm_qkv -f python bench.py --profile-run > $OUT/ncu_gemm_qkv.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_bf16_tn_kernelILi128ELi6ELi0ELb0ELb0 -s 60 -c 2 -o $OUT/prof_gemm_out_ff2 -f python bench.py --profile-run > $OUT/ncu_gemm_out.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_bf16_tn_kernelILi128ELi3ELi1ELb1ELb0 -s 30 -c 1 -o $OUT/prof_gemm_ff1 -f python bench.py --profile-run > $OUT/ncu_gemm_ff1.log 2>&1
ncu --set full --clock-control none --cache-control none --kernel-name-base mangled -k "regex:gemm_bf16_tn_kernelILi128ELi6ELi0ELb0ELb0|gemm2_bf16_tn_kernelILi192" -s 90 -c 3 -o $OUT/prof_gemm_warm -f python bench.py --profile-run > $OUT/ncu_gemm_warm.log 2>&1
ls -la $OUT/prof_gemm*.ncu-rep
