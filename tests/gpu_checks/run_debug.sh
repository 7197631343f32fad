export PYTHONPATH=.
F5_ATTN_VARIANT=3 python tests/gpu_checks/check_attention.py 2>&1 | tail -12
F5_ATTN_VARIANT=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B1 attn v3: ms/step', d['ms_per_step'], 'value', d['value'])"
