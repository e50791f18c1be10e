#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ecapa.py tests/test_gpu_fbank.py tests/test_gpu_e2e.py -q -x --timeout=600 > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_g.log
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
python tools/ab_probe.py cur= cur2=
