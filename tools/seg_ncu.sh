#!/bin/bash
O=gpurun_out; mkdir -p $O
export LNA_ESS_SEG=1 LNA_ESS_SEG_T=32 LNA_GRAPH=0
python tools/eslice_profile.py 65536 3 > $O/seg_plain.out 2>&1 || { echo "plain failed"; tail -5 $O/seg_plain.out; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:k_ess_par_seg' -s 1 -c 1 -o $O/r2_seg python tools/eslice_profile.py 65536 2 > $O/ncu_seg.out 2>&1
ncu -i $O/r2_seg.ncu-rep --page raw --csv > $O/r2_seg_raw.csv 2>/dev/null
ncu -i $O/r2_seg.ncu-rep --page source --csv > $O/r2_seg_source.csv 2>/dev/null
ls -la $O/r2_seg*
