#!/bin/bash
# HEAD validation: gpu tests, bench (+reference arm), launch list, ncu full of the fused kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
# launch list of the bench command itself (a short run: 3 warm-up + 2 timed steps, the per-stage pass, the e2e leg)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ahc > gpurun_out/ncu_launches.log 2>&1
echo "bench launch list rc=$?"
# and of the capture target (one forward pair + affinity / AHC + the rank-4 operators)
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_target.csv python tools/ncu_target.py > gpurun_out/ncu_launches_target.log 2>&1
echo "target launch list rc=$?"
if [ "$1" == "full" ]; then bash tools/ncu_gemm.sh mfa tdnn2 r2f att pool fbank se aff ahc post; fi
if [ "$1" == "some" ]; then shift; bash tools/ncu_gemm.sh "$@"; fi
