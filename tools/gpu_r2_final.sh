#!/bin/bash
# round 2 final evidence on one GPU: tests, bench (both arms + library arm), launch list of the bench itself
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --impl torch_gpu --steps 10 --warmup 3 > gpurun_out/bench_torch_gpu.json 2>> gpurun_out/bench.err; echo "torch rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "clocks", "gpu_launches")})
print("e2e", json.dumps(d["e2e"]["legs"]))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "launch_ms")}, d["roofline"]["whole_step"])
print("stage", json.dumps(d["roofline"]["stage_ms"]))
print("cpu", json.dumps(d.get("cpu_baseline")))
lb = d.get("library_baseline") or {}
print("lib", {k: lb.get(k) for k in ("value", "ms_per_step", "speedup_value", "speedup_e2e_vs_reference_call")}, (lb.get("e2e") or {}).get("value"))
print("dense", json.dumps(d.get("dense_pass")))
print("ahc", json.dumps({k: {kk: v.get(kk) for kk in ("affinity_ms", "ahc_ms", "labels_match_planted")} for k, v in d["ahc"].items() if isinstance(v, dict)}))
print("post", json.dumps(d.get("post"))[:600])
PY
cat gpurun_out/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ahc --no-library-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
