#!/bin/bash
# ncu evidence of round 2 (GPU box, repo root).  Every profiled command first runs without ncu.
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eslice-iters 2 --no-extra"
$B > $O/ncu_plain.json 2> $O/ncu_plain.err || { echo "plain bench failed"; tail -5 $O/ncu_plain.err; exit 1; }
LNA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2_launches_bench_ncu.csv $B > $O/ncu_list.out 2>&1
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
$N -k 'regex:k_full_loglik<\(int\)0, \(bool\)0' -s 12 -c 1 -o $O/r2_full $B --no-eslice --no-p3 > $O/ncu_full.out 2>&1
$N -k 'regex:k_full_loglik<\(int\)1' -s 6 -c 1 -o $O/r2_seir $B --no-eslice > $O/ncu_seir.out 2>&1
$N -k 'regex:k_sirs_full' -s 6 -c 1 -o $O/r2_sirs $B --no-eslice > $O/ncu_sirs.out 2>&1
python tools/eslice_profile.py 4096 6 > $O/ncu_iter_plain.out 2>&1 || { echo "plain iteration failed"; exit 1; }
LNA_GRAPH=0 $N -k 'regex:k_cand_eval|k_eslice_traj' -s 15 -c 4 -o $O/r2_iter python tools/eslice_profile.py 4096 6 > $O/ncu_iter.out 2>&1
for r in r2_full r2_seir r2_sirs r2_iter; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ls -la $O/*.ncu-rep $O/*_raw.csv
