export PYTHONPATH=.
timeout 300 python tests/gpu_checks/check_attention.py 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_b1.json; python -c "import json; d=json.load(open('gpurun_out/bench_b1.json')); print('B1: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
timeout 600 python bench.py --steps 3 --warmup 3 --batch 64 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_b64.json; python -c "import json; d=json.load(open('gpurun_out/bench_b64.json')); print('B64: ms/step', d['ms_per_step'], 'value', d['value'])"
