#!/bin/bash
# ncu --set full of the barrier-free slice kernel (GPU box)
O=gpurun_out; mkdir -p $O
python tools/eslice_profile.py 4096 6 > $O/async_plain.out 2>&1 || { echo "plain run failed"; tail -5 $O/async_plain.out; exit 1; }
LNA_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:k_ess_par_async' -s 3 -c 1 \
  -o $O/r2_async python tools/eslice_profile.py 4096 6 > $O/ncu_async.out 2>&1
ncu -i $O/r2_async.ncu-rep --page raw --csv > $O/r2_async_raw.csv 2>/dev/null
ncu -i $O/r2_async.ncu-rep --page source --csv > $O/r2_async_source.csv 2>/dev/null
ls -la $O/r2_async*
