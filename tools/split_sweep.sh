#!/bin/bash
# Experiment (GPU box): three-warp pipeline threshold x slice phase plan for the 4096-chain iteration (one stream group, graph replay)
B=${1:-4096}
for W in "" 592 1184 2368; do
  for PLAN in "" "8:0:1,8:8:1,8:16:0" "8:0:1,4:8:1,16:12:0" "8:0:1,4:8:1,4:12:1,8:16:0"; do
    for PAIR in "" 0; do
      env ${W:+LNA_SPLIT_MAX_WARPS=$W} ${PLAN:+LNA_SLICE_PLAN=$PLAN} ${PAIR:+LNA_MH_PAIR=$PAIR} LNA_CHAIN_GROUPS=1 \
        python tools/plan_sweep.py child $B 2>&1 | tail -1 | sed "s/^/split_max_warps=${W:-default} plan=${PLAN:-auto} mh_pair=${PAIR:-auto}  /"
    done
  done
done
