#!/bin/bash
# One GPU session: tests, bench, ncu launch list, ncu full captures.  Run under gpurun.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
if [ "$1" == "full" ]; then
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<1, 256>" -s 15 -c 1 -o gpurun_out/prof_mfa python tools/ncu_target.py > gpurun_out/ncu_mfa.log 2>&1; echo "mfa rc=$?"
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<1, 128>" -s 25 -c 1 -o gpurun_out/prof_res2net python tools/ncu_target.py > gpurun_out/ncu_res.log 2>&1; echo "res rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:fbank_frames -s 1 -c 1 -o gpurun_out/prof_fbank python tools/ncu_target.py > gpurun_out/ncu_fbank.log 2>&1; echo "fbank rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:ahc_rounds -c 1 -o gpurun_out/prof_ahc python tools/ncu_target.py > gpurun_out/ncu_ahc.log 2>&1; echo "ahc rc=$?"
fi
