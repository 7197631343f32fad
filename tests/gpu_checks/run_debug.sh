export PYTHONPATH=.
for pdl in 0 1; do
F5_PDL=$pdl python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
done
F5_PDL=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
F5_PDL=1 python bench.py --steps 3 --warmup 3 --batch 64 --method midpoint --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B64 midpoint PDL=1: ms/step', d['ms_per_step'], 'value', d['value'], d['roofline']['achieved'], d['roofline']['whole_step'])"
python bench.py --steps 3 --warmup 3 --batch 64 --method midpoint --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B64 midpoint PDL=0: ms/step', d['ms_per_step'], 'value', d['value'], d['roofline']['achieved'], d['roofline']['whole_step'])"
