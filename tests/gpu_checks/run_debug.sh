export PYTHONPATH=.
for po in 0 1 2; do echo "== poly=$po"; F5_ATTN_POLY=$po timeout 300 python tests/gpu_checks/check_attention.py 2>&1 | tail -8; done
for po in 0 1; do F5_ATTN_POLY=$po timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1 poly=$po: ms/step', d['ms_per_step'], 'value', d['value'])"; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
