#!/bin/bash
mkdir -p gpurun_out
for N in 5000 20000 50000; do python tools/ahc_time.py $N; SD_AHC_F32=1 python tools/ahc_time.py $N; done
SD_AHC_F32=1 timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_centroid.py tests/test_gpu_e2e.py -q --timeout=600 2>&1 | tail -12
