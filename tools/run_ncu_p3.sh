O=gpurun_out; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eslice-iters 2 --no-extra"
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
$N -k 'regex:k_full_loglik<\(int\)0, \(bool\)0' -s 12 -c 1 -o $O/r2_full $B --no-eslice --no-p3 > $O/ncu_full.out 2>&1
$N -k 'regex:k_full_loglik<\(int\)1' -s 6 -c 1 -o $O/r2_seir $B --no-eslice > $O/ncu_seir.out 2>&1
for r in r2_full r2_seir; do [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; done
ls -la $O/r2_full* $O/r2_seir*
python -m pytest tests -m gpu -q -x -k "sirs or selfmix or self_mixing" 2>&1 | tail -5
