export PYTHONPATH=.
timeout 600 python tests/gpu_checks/check_gemm.py 2>&1 | grep -v "^OK" | tail -4
timeout 600 python tests/gpu_checks/check_dit.py 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1: ms/step', d['ms_per_step'], 'value', d['value'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
