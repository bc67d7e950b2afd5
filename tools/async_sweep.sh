#!/bin/bash
# Experiment (GPU box): the barrier-free early-rejection slice kernel (lna_async.cuh) against the wave kernels
for B in ${@:-4096}; do
  LNA_ESS_ASYNC=0 python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/B=$B wave kernels          /"
  for W in 4 8 16; do
    for WPS in 1 2 3; do
      LNA_ESS_ASYNC_W=$W LNA_ESS_ASYNC_WPS=$WPS python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/B=$B async W=$W wps=$WPS       /"
    done
  done
  python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/B=$B default               /"
done
