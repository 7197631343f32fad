export PYTHONPATH=.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_b1.err | tail -1 > gpurun_out/bench_final_b1.json
python -c "import json; d=json.load(open('gpurun_out/bench_final_b1.json')); print('B1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
