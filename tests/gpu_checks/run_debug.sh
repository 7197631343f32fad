# scratch driver for bring-up runs on the GPU box:  gpurun -- 'bash tests/gpu_checks/run_debug.sh'
export PYTHONPATH=.
timeout 300 python tests/gpu_checks/check_gemm.py 2>&1 | tail -12
timeout 300 python tests/gpu_checks/check_attention.py 2>&1 | tail -12
timeout 120 python tests/gpu_checks/check_attn_timeline.py 2 937 16
