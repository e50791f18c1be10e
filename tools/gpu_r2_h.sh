#!/bin/bash
# fbank round: full GPU pytest, bench line, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],{k:v['value'] for k,v in d['e2e']['legs'].items()})
print(d['roofline']['stage_ms']); print(d['clocks']); print(d.get('library_baseline',{}).get('f16_autocast'), d.get('library_baseline',{}).get('value'))
PY
