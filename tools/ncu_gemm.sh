#!/bin/bash
# ncu --set full captures (second forward of tools/ncu_target.py).
# gemm_tc_2sm launches per forward: block0, (tdnn1, tdnn2) x3, mfa  -> second forward = indices 8..15
# usage: tools/ncu_gemm.sh [mfa] [tdnn2] [r2p] [att] [pool] [fbank] [se] [ahc] [aff] [post]
set -x
mkdir -p gpurun_out
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
N='--set full --clock-control none --import-source on --kernel-name-base demangled'
for what in "$@"; do
case $what in
 mfa)   ncu $N -k 'regex:gemm_tc_2sm_kernel' -s 15 -c 1 -o gpurun_out/prof_mfa python tools/ncu_target.py > gpurun_out/ncu_mfa.log 2>&1;;
 tdnn2) ncu $N -k 'regex:gemm_tc_2sm_kernel' -s 10 -c 1 -o gpurun_out/prof_tdnn2 python tools/ncu_target.py > gpurun_out/ncu_tdnn2.log 2>&1;;
 r2p)   ncu $N -k 'regex:res2net_pipe' -s 4 -c 1 -o gpurun_out/prof_r2p python tools/ncu_target.py > gpurun_out/ncu_r2p.log 2>&1;;
 att)   ncu $N -k 'regex:gemm_tc_kernel<\(int\)4, \(int\)128>' -s 1 -c 1 -o gpurun_out/prof_att python tools/ncu_target.py > gpurun_out/ncu_att.log 2>&1;;
 pool)  ncu $N -k 'regex:gemm_tc_kernel<\(int\)2, \(int\)256>' -s 1 -c 1 -o gpurun_out/prof_pool python tools/ncu_target.py > gpurun_out/ncu_pool.log 2>&1;;
 fbank) ncu $N -k 'regex:fbank_frames' -s 1 -c 1 -o gpurun_out/prof_fbank python tools/ncu_target.py > gpurun_out/ncu_fbank.log 2>&1;;
 ahc)   ncu $N -k 'regex:ahc_rounds' -c 1 -o gpurun_out/prof_ahc python tools/ncu_target.py > gpurun_out/ncu_ahc.log 2>&1;;
 aff)   ncu $N -k 'regex:gemm_tc_kernel<\(int\)3' -c 1 -o gpurun_out/prof_aff python tools/ncu_target.py > gpurun_out/ncu_aff.log 2>&1;;
 se)    ncu $N -k 'regex:se_gate|se_apply|colstats_finish' -s 7 -c 3 -o gpurun_out/prof_se python tools/ncu_target.py > gpurun_out/ncu_se.log 2>&1;;
 post)  ncu $N -k 'regex:viterbi_forward|cohort_topk|jacobi_kernel|gram_kernel|hysteresis|mask_segments' -c 7 -o gpurun_out/prof_post python tools/ncu_target.py > gpurun_out/ncu_post.log 2>&1;;
esac
echo "$what rc=$?"
done
ls -la gpurun_out/*.ncu-rep
