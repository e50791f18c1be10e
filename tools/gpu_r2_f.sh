#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:res2net_pipe' -s 4 -c 1 -o gpurun_out/prof_r2p python tools/ncu_target.py > gpurun_out/ncu_r2p.log 2>&1; echo "r2p rc=$?"
ls -la gpurun_out/*.ncu-rep
