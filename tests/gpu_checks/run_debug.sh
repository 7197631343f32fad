export PYTHONPATH=.
python tests/gpu_checks/run.py check_gemm check_gemm2 2>&1 | grep -E "OK|FAIL|time M1874|ALL" | cut -c 1-330
python tests/gpu_checks/check_insitu.py 2>&1 | tail -14
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -2 gpurun_out/bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench4.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d['roofline'])[:900])
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
