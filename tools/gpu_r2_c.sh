#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 4 7; do
  echo "== DBG $v (cols: mma_start mma_issued | wait t_full done arrive)"
  SD_LIB_PATH=$PWD/speech_diarization_b200/csrc/build_v/trace$v.so python tools/r2p_trace.py 2>&1 | sed -n '10,15p' | cut -c1-75
done
