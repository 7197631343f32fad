export PYTHONPATH=.
timeout 300 python tests/gpu_checks/check_insitu.py 2>&1 | tail -28
