#!/bin/bash
# Experiment (GPU box): barrier-free slice kernel with dynamic candidate assignment -- lanes per chain x resident warps per
# sub-partition x batch size, against the wave kernels
for B in ${@:-4096}; do
  LNA_ESS_ASYNC=0 python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/B=$B wave kernels      /"
  for cfg in "8 2" "8 3" "16 3" "16 4"; do
    set -- $cfg
    LNA_ESS_ASYNC_W=$1 LNA_ESS_ASYNC_WPS=$2 python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/B=$B async W=$1 wps=$2   /"
  done
done
