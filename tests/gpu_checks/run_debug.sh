export PYTHONPATH=.
python tests/gpu_checks/check_insitu.py 2>&1 | tail -16
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
