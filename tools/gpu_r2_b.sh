#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ecapa.py -q -x --timeout=600 -k "fused_res2net or time_statistics or slot or overflow or large_batchnorm or pageable or batch_properties" > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r2.log
python tools/ab_probe.py pipe= mc=,SD_R2_MC=1 v1=,SD_R2_PIPE=0 pipe2=
