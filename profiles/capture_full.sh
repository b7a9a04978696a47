#!/bin/bash
# ncu --set full capture (with source) of the record kernels of one resident pass.  usage (on the GPU box): profiles/capture_full.sh <tag> <shape> [MB]
# The pass runs with one subblock group (PHY_GROUPS=1), so every kernel appears once per pass; the first pass is skipped.
TAG=$1; SHAPE=${2:-36bp}; MB=${3:-256}
K='regex:k_nl_count|k_nl_emit|k_plan|k_stat1|k_seqstat|k_stat2|k_enc_title|k_enc_qd|k_place|k_classify|k_huff'
N=11
PHY_GROUPS=1 timeout 900 ncu --set full --import-source on --clock-control none --kernel-name "$K" --launch-skip $N --launch-count $N \
  -f -o gpurun_out/${TAG}_full_${SHAPE} python tests/gpu_prof_target.py $SHAPE $MB 2 > gpurun_out/${TAG}_ncu_${SHAPE}.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_${SHAPE}.log
