#!/bin/bash
# r02 call 7 (2 GPUs): NCCL sharded==unsharded + C-ABI broadcast, 2-GPU bench (incl. 64/GPU sub-result), reference arm under torchrun
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 500 -k nccl > $OUT/r02_c7_nccl.log 2>&1; tail -5 $OUT/r02_c7_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/gpu_checks/nccl_shard_check.py 2>&1 | grep -E "NCCL_SHARD_CHECK|Error|error" | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 5 --warmup 3 2> $OUT/r02_c7_bench2.err | tail -1 > $OUT/r02_c7_bench_2gpu.json
tail -c 300 $OUT/r02_c7_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29743 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2> $OUT/r02_c7_ref2.err | tail -1 > $OUT/r02_c7_bench_ref_2gpu.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_c7_bench_2gpu.json"))
print("2gpu ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"], d["config"]["global_batch"])
for k, v in d["configs"].items():
    print("  ", k, v["ms_per_step"], v["value"], v["config"]["global_batch"])
r = json.load(open("gpurun_out/r02_c7_bench_ref_2gpu.json"))
print("ref", r["value"], r["config"]["global_batch"], r["steps_run"], r["config"] == d["config"])
PY
