#!/bin/bash
# compute-sanitizer evidence for the pipelined (three-warp, shared-memory ring) kernels and the one-warp kernels.
# usage (GPU box, repo root): bash tools/run_sanitizers.sh   -> gpurun_out/r2_{memcheck,racecheck,synccheck}.log
set -u
mkdir -p gpurun_out
python tools/sanitize_probe.py 2 || exit 1
compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck.log python tools/sanitize_probe.py 1 > gpurun_out/r2_memcheck.out 2>&1
tail -3 gpurun_out/r2_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/r2_racecheck.log python tools/sanitize_probe.py 1 > gpurun_out/r2_racecheck.out 2>&1
tail -5 gpurun_out/r2_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --log-file gpurun_out/r2_synccheck.log python tools/sanitize_probe.py 1 > gpurun_out/r2_synccheck.out 2>&1
tail -3 gpurun_out/r2_synccheck.log
