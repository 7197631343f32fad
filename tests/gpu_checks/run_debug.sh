export PYTHONPATH=.
python tests/gpu_checks/check_determinism.py 2>&1 | tail -8
python tests/gpu_checks/check_cfg_equiv.py 2>&1 | tail -8
python tests/gpu_checks/check_attention.py 2>&1 | tail -6
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 1800 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
F5_PDL=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL off: ms/step', d['ms_per_step'], 'value', d['value'])"
