#!/bin/bash
# round 2, multi-GPU session: the NCCL tests and the scaling bench with the sharded clustering leg
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout=600 -rA > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "pytest multi rc=$?"
tail -15 gpurun_out/pytest_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
print("cluster", json.dumps(d.get("cluster")))
PY
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu.json"))
print("corpus", json.dumps(d.get("corpus_8h")))
PY
