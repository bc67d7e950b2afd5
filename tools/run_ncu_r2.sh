#!/bin/bash
# Round-2 evidence (GPU box, repo root): tests, bench lines, ncu launch list and --set full captures.  Every profiled command
# first runs without ncu.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_final_pytest.log
python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err || { echo "bench failed"; tail -5 $O/r2_bench_1gpu.err; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eslice-iters 2 --no-extra"
$B > $O/ncu_plain.json 2> $O/ncu_plain.err || { echo "plain bench failed"; tail -5 $O/ncu_plain.err; exit 1; }
LNA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_bench_ncu.csv $B > $O/ncu_list.out 2>&1
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
# the headline launch: B = 65536 = 1024 blocks of 64 threads (launches 100+ of this kernel sit in the bench's pre-load loop)
$N -k 'regex:k_full_loglik<\(int\)0, \(bool\)0' -s 100 -c 1 -o $O/r2_full $B --no-eslice --no-p3 > $O/ncu_full.out 2>&1
$N -k 'regex:k_full_loglik<\(int\)1' -s 6 -c 1 -o $O/r2_seir $B --no-eslice > $O/ncu_seir.out 2>&1
$N -k 'regex:k_sirs_full' -s 6 -c 1 -o $O/r2_sirs $B --no-eslice > $O/ncu_sirs.out 2>&1
python tools/eslice_profile.py 4096 6 > $O/ncu_iter_plain.out 2>&1 || { echo "plain iteration failed"; exit 1; }
LNA_GRAPH=0 $N -k 'regex:k_cand_eval|k_eslice_traj|k_ess_par_async' -s 12 -c 3 -o $O/r2_iter python tools/eslice_profile.py 4096 6 > $O/ncu_iter.out 2>&1
for r in r2_full r2_seir r2_sirs r2_iter; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ls -la $O/*.ncu-rep $O/*_raw.csv; cat $O/r2_final_pytest.log
