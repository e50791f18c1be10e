#!/bin/bash
# round 2, GPU session A: all GPU tests, the bench line with the new legs, the GPU-library arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "clocks", "gpu_launches")})
    print("e2e", json.dumps(d["e2e"]))
    print("stage", json.dumps(d["roofline"]["stage_ms"]))
    print("lib", json.dumps(d.get("library_baseline")))
    print("ahc", json.dumps(d.get("ahc")))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl torch_gpu --steps 10 --warmup 3 > gpurun_out/bench_torch_gpu.json 2> gpurun_out/bench_torch_gpu.err; echo "torch_gpu rc=$?"
cat gpurun_out/bench_torch_gpu.json
