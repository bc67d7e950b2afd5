#!/bin/bash
# Round-2 (second half) evidence (GPU box, repo root): bench lines, ncu launch list and --set full captures of the kernels that
# changed -- the balanced FULL kernel (headline launch) and the barrier-free slice kernel with dynamic candidate assignment.
# Every profiled command first runs without ncu.
set -u
O=gpurun_out; mkdir -p $O
python bench.py > $O/r2b_bench_1gpu.json 2> $O/r2b_bench_1gpu.err || { echo "bench failed"; tail -5 $O/r2b_bench_1gpu.err; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2b_bench_reference.json 2> $O/r2b_bench_reference.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eslice-iters 2 --no-extra"
$B > $O/ncu_plain.json 2> $O/ncu_plain.err || { echo "plain bench failed"; tail -5 $O/ncu_plain.err; exit 1; }
LNA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2b_launches_bench_ncu.csv $B > $O/ncu_list.out 2>&1
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
# the headline launch: B = 65536 on 148 persistent blocks of 512 threads
python tools/headline_kernel.py 6 > $O/headline_plain.out 2>&1 || { echo "headline failed"; exit 1; }
$N -k 'regex:k_full_loglik_bal' -s 3 -c 1 -o $O/r2b_full python tools/headline_kernel.py 6 > $O/ncu_full.out 2>&1
python tools/eslice_profile.py 4096 6 > $O/ncu_iter_plain.out 2>&1 || { echo "plain iteration failed"; exit 1; }
LNA_GRAPH=0 $N -k 'regex:k_ess_par_async' -s 3 -c 1 -o $O/r2b_async8 python tools/eslice_profile.py 4096 6 > $O/ncu_async8.out 2>&1
LNA_GRAPH=0 $N -k 'regex:k_ess_par_async' -s 3 -c 1 -o $O/r2b_async16 python tools/eslice_profile.py 2048 6 > $O/ncu_async16.out 2>&1
for r in r2b_full r2b_async8 r2b_async16; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
[ -f $O/r2b_full.ncu-rep ] && ncu -i $O/r2b_full.ncu-rep --page source --csv > $O/r2b_full_source.csv 2>/dev/null
[ -f $O/r2b_async8.ncu-rep ] && ncu -i $O/r2b_async8.ncu-rep --page source --csv > $O/r2b_async8_source.csv 2>/dev/null
LNA_TIMELINE=1 python tools/async_timeline.py 4096 > $O/r2b_timeline_4096.log 2>&1
ls -la $O/r2b_*.ncu-rep $O/r2b_*_raw.csv
