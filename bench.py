#!/usr/bin/env python
"""bench.py -- LNA+coalescent log-likelihood evaluations/sec (and ESlice chain-iterations/sec).

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1) prints ONE JSON
line on rank 0.  Workload = BASELINE.json configs[1] (Changpoint_sim vignette: SIR LNA, N=1e6,
2000 fine steps, gridsize 50, 39 R0 change points, the vignette's cached 342-tip genealogy E=376),
4096 chains.  One "step" of the likelihood leg = one batch of B FULL evaluations
(param, initial, OriginTraj) -> coalLog per GPU, B = 4096 chains x 16 slice proposals by default.

  value      : FULL evaluations/s, inputs already resident in HBM, CUDA-event timed on the launch stream
  e2e        : the same through the host-buffer C ABI call (pinned host inputs, H2D + kernel + D2H)
  eslice     : chain-iterations/s of the batched General_MCMC_with_ESlice iteration (when built)
  roofline   : algorithmic FP64 flops (113/fine step + 21/interval + 10/cell)
               over the measured independent-DFMA peak of this GPU
  cpu_baseline: the reference's CPU implementation on the host cores, bounded sample: the reference's own
               sources compiled verbatim on stand-in Armadillo/Rcpp headers (oracle/_ref, kind "reference") when
               that library is present, else the C restatement (kind "port")

`--impl reference` times only that CPU arm (R/Rcpp/Armadillo are absent, so the reference's package build cannot
run; its C++ sources compile as they are against oracle/shim/).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

_JSON_OUT = sys.stdout
METRIC = "LNA+coalescent loglik evals/sec (FULL evaluations, batched chains)"
UNIT = "evals/s"
WORKLOAD = ("configs[1] Changpoint_sim vignette: SIR LNA N=1e6, 2000 fine steps, gridsize 50, 39 R0 change points, "
            "cached 342-tip genealogy (E=376); FULL (param, initial, OriginTraj) -> coalLog evaluations")
# DRAM bytes per launch of the hot kernel at the default batch (dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture, profiles/r1d_ncu_full_loglik.md)
NCU_TRAFFIC_BYTES_65536 = 65061120 + 1034752


def load_workload():
    from conftest import Golden
    return Golden("changepoint_sim")


def synth_inputs(g, B, seed):
    """Synthetic draws of the vignette's shape: parameters around the cached chain state, fresh
    N(0,1) latent increments.  Natural layouts: par (B,44), eps (B,k,p)."""
    rng = np.random.default_rng(seed)
    par = np.tile(g.final["par"].ravel(), (B, 1))
    par[:, 2] *= np.exp(0.05 * rng.standard_normal(B))
    par[:, 3] *= np.exp(0.05 * rng.standard_normal(B))
    par[:, 4:43] *= np.exp(0.02 * rng.standard_normal((B, 39)))
    # latent increments like slice proposals around the cached state: eps*cos(th) + N(0,1)*sin(th)
    th = rng.uniform(-0.3, 0.3, size=(B, 1, 1))
    eps = g.final["OriginTraj"][None] * np.cos(th) + rng.standard_normal((B, 40, 2)) * np.sin(th)
    return par, eps


def algorithmic_flops(g):
    n = len(g.times) - 1
    k = n // g.gridsize
    ng = int(g.init["ng"][0])
    return 113.0 * n + 21.0 * k + 10.0 * (ng - 1)


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_impl():
    """The CPU arm: the reference's own sources compiled verbatim (oracle/_ref/liblna_ref.so, kind "reference") when that
    library is here, else the C restatement (kind "port").  Both live under oracle/ and are only ever timed/checked."""
    from oracle import reference as ref
    if ref.available():
        return ref, "reference", ("reference code as written: /root/reference/src/LNA_functional.cpp (+ basic_util, SIR_LNA, "
                                  "SIRS_period, SIR_phylodyn) compiled verbatim, g++ -O2, on the Armadillo/Rcpp stand-in "
                                  "headers of oracle/shim/; each evaluation is one Update_Param call")
    from oracle import oracle as orc
    return orc, "port", "CPU oracle = allocation-free C restatement of the reference algorithm, gcc -O2 -ffp-contract=off"


def time_oracle(g, n_per_thread, threads, seed=123, impl=None):
    if impl is None:
        from oracle import oracle as impl
    orc = impl
    cfg, ini = orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init)
    par, eps = synth_inputs(g, n_per_thread * threads, seed)
    chunks = [(par[i * n_per_thread:(i + 1) * n_per_thread], eps[i * n_per_thread:(i + 1) * n_per_thread])
              for i in range(threads)]
    outs = [None] * threads

    def work(i):
        outs[i] = orc.full_loglik_batch(cfg, ini, chunks[i][0][:, 2:], chunks[i][0][:, :2], g.times, g.gridsize,
                                        chunks[i][1], g.t_correct)

    orc.lib()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return n_per_thread * threads / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = load_workload()
    cores = os.cpu_count() or 1
    impl, kind, what = cpu_impl()
    per = args.ref_evals_per_core if args.ref_evals_per_core > 0 else (96 if kind == "reference" else 1024)
    per = max(16, int(per))
    for _ in range(args.warmup):
        time_oracle(g, 16, cores, impl=impl)
    tot_t, tot_n = 0.0, 0
    for s in range(args.steps):
        rate, dt = time_oracle(g, per, cores, seed=1000 + s, impl=impl)
        tot_t += dt
        tot_n += per * cores
    val = tot_n / tot_t
    sample = f"{per} FULL evals per core per step x {cores} threads (Changpoint_sim shape); {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_evals_per_step": per * cores},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def cpu_reference(g, iters=12):
    """bench.py's CPU leg for the ESlice metric: the restated reference driver (test infrastructure under
    oracle/, timed only -- never on the product path) on ONE host core, a bounded sample of chain-iterations."""
    import time
    from lnaphylodyn_b200.chains import ITER_UNIF
    impl, kind, what = cpu_impl()
    if kind == "reference":
        driver = impl.driver()
    else:
        from oracle import driver
    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]],
                           g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    st = driver.init_state(g, par0, g.setting("control.traj"))
    opts_b, opts = driver.opts_from_golden(g, update_traj=False), driver.opts_from_golden(g, update_traj=True)
    rng = np.random.default_rng(99)
    for _ in range(4):
        st = driver.eslice_iteration(st, g, opts_b, rng.standard_normal(123), rng.random(ITER_UNIF))
    n0 = st["n_full"]
    t0 = time.perf_counter()
    for _ in range(iters):
        st = driver.eslice_iteration(st, g, opts, rng.standard_normal(123), rng.random(ITER_UNIF))
    dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": "chain-iterations/s", "cores": 1, "kind": kind,
            "sample": f"{iters} iterations of one chain, the R-level iteration body (oracle/driver.py) over: {what}"}


def bench_eslice(pb, g, B, iters, rank, world, dev):
    """ESlice chain-iterations/s: B chains per GPU from the vignette's initial state, streams resident
    in HBM (device-timed) and through host buffers (e2e)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from lnaphylodyn_b200 import _lib
    from lnaphylodyn_b200._lib import check
    from lnaphylodyn_b200.chains import Chains, ITER_UNIF, vignette_opts

    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]],
                           g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    par = np.tile(par0, (B, 1))
    eps0 = g.setting("control.traj")
    ch = Chains(pb, par, eps0)
    opts_burn = vignette_opts(g, update_traj=False)
    opts = vignette_opts(g, update_traj=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    nsets = iters + 8
    nrm = torch.randn((nsets, B, ch.n_norm), dtype=torch.float64, device=dev, generator=gen)
    unf = torch.rand((nsets, B, ITER_UNIF), dtype=torch.float64, device=dev, generator=gen).clamp_(1e-300, 1 - 1e-16)
    stream = torch.cuda.current_stream(dev)
    # burn-in like the vignette's first stage, then warm-up of the full iteration
    for i in range(4):
        ch.step_dev(opts_burn, nrm[i].data_ptr(), unf[i].data_ptr())
    for i in range(4, 8):
        ch.step_dev(opts, nrm[i].data_ptr(), unf[i].data_ptr())
    torch.cuda.synchronize(dev)
    before = ch.get()
    if world > 1:
        dist.barrier()
    l0 = pb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        ch.step_dev(opts, nrm[8 + i].data_ptr(), unf[8 + i].data_ptr())
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launches = pb.launch_count() - l0
    after = ch.get()
    # e2e: host streams (pinned) -> H2D -> iteration -> D2H of the summary block
    nrm_h = torch.randn((B, ch.n_norm), dtype=torch.float64).pin_memory()
    unf_h = torch.rand((B, ITER_UNIF), dtype=torch.float64).clamp_(1e-300, 1 - 1e-16).pin_memory()
    summ_h = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    L = _lib.lib()
    cl_h = torch.empty(B, dtype=torch.float64).pin_memory()
    lo_h = torch.empty(B, dtype=torch.float64).pin_memory()
    for i in range(3):  # first-touch of the pinned buffers and of the host-pointer staging path, untimed
        check(L.lna_chains_step(ch.h, C.byref(opts), nrm_h.data_ptr(), unf_h.data_ptr()), "lna_chains_step")
        check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
    t0 = time.perf_counter()
    for i in range(iters):
        check(L.lna_chains_step(ch.h, C.byref(opts), nrm_h.data_ptr(), unf_h.data_ptr()), "lna_chains_step")
        check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
    e2e_s = time.perf_counter() - t0
    # same end-to-end loop with the streams drawn on the device (Philox): only the summaries cross PCIe
    t0 = time.perf_counter()
    for i in range(iters):
        check(L.lna_chains_step_rng(ch.h, C.byref(opts), 20261020, 1000 + i, rank * B), "lna_chains_step_rng")
        check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
    rng_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s * 1e3, rng_s * 1e3], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float((after["n_full_evals"] - before["n_full_evals"]).sum()),
                        float((after["n_cheap_evals"] - before["n_cheap_evals"]).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms, rng_ms = float(t[0]), float(t[1]), float(t[2])
    tot = world * B * iters
    # the one exchange of the path: all-gather of the per-chain summary rows, NCCL over NVLink
    gather_us, gathered_rows = None, None
    if world > 1:
        from lnaphylodyn_b200.chains import gather_summaries

        class _Dev:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (int(ptr), False), "version": 2}

        local = torch.as_tensor(_Dev(ch.summary_ptr(), (B, 4)), device=dev)
        for _ in range(3):
            allrows = gather_summaries(local)
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(10):
            allrows = gather_summaries(local)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        tg = torch.tensor([g0.elapsed_time(g1) * 100.0], dtype=torch.float64, device=dev)  # us per gather
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_us, gathered_rows = float(tg[0]), int(allrows.shape[0])
        assert gathered_rows == world * B and bool(torch.isfinite(allrows[:, 0]).all())
    cpu = None
    if rank == 0:
        cpu = cpu_reference(g)
    return {"value": tot / (ms * 1e-3), "unit": "chain-iterations/s", "chains_per_gpu": B, "iters": iters,
            "ms_per_iter": ms / iters, "e2e_value": tot / (e2e_ms * 1e-3),
            "e2e_device_rng_value": tot / (rng_ms * 1e-3),
            "h2d_bytes_per_iter": int(nrm_h.numel() * 8 + unf_h.numel() * 8), "d2h_bytes_per_iter": int(B * 16),
            "mean_full_evals_per_iter": float(cnt[0]) / tot, "mean_cheap_evals_per_iter": float(cnt[1]) / tot,
            "consumed_full_evals_per_s": float(cnt[0]) / (ms * 1e-3), "gpu_launches": int(launches),
            "spec_width": "auto", "finite_coalLog": bool(np.all(np.isfinite(after["coalLog"]))),
            "summary_allgather_us": gather_us, "summary_rows_gathered": gathered_rows,
            "cpu_baseline": cpu}



# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=4096, help="concurrent chains per GPU")
    ap.add_argument("--proposals", type=int, default=16, help="slice proposals per chain in one likelihood batch")
    ap.add_argument("--ref-evals-per-core", type=int, default=0, help="0 = 96 (reference as written) / 1024 (port)")
    ap.add_argument("--cpu-evals-per-core", type=int, default=0, help="0 = 512 (reference as written) / 4096 (port)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eslice", action="store_true")
    ap.add_argument("--no-seir", action="store_true")
    ap.add_argument("--eslice-iters", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 (NCCL prints its version banner
    # there when NCCL_DEBUG is set on the box) are sent to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from lnaphylodyn_b200 import _lib, api

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    L = _lib.lib()
    g = load_workload()
    B = args.chains * args.proposals
    pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    _lib.check(L.lna_problem_set_stream(pb.h, C.c_void_p(stream.cuda_stream)), "set_stream")

    # ---- inputs: NBUF distinct batches resident in HBM (this rank's shard of the global batch) ----
    NBUF = 4
    bufs = []
    host = []
    for i in range(NBUF):
        par, eps = synth_inputs(g, B, seed=10_000 * (rank + 1) + i)
        param_h = torch.from_numpy(np.ascontiguousarray(par[:, 2:])).pin_memory()
        init_h = torch.from_numpy(np.ascontiguousarray(par[:, :2])).pin_memory()
        eps_h = torch.from_numpy(np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))).pin_memory()  # item = k x p col-major
        host.append((param_h, init_h, eps_h))
        bufs.append(tuple(t.to(dev) for t in (param_h, init_h, eps_h)))
    coal_d = torch.empty(B, dtype=torch.float64, device=dev)
    st_d = torch.empty(B, dtype=torch.int32, device=dev)
    coal_h = torch.empty(B, dtype=torch.float64).pin_memory()
    st_h = torch.empty(B, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def step_dev(i):
        p_, i_, e_ = bufs[i % NBUF]
        _lib.check(L.lna_full_loglik_dev(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), coal_d.data_ptr(),
                                         st_d.data_ptr(), None, None, None, None, None), "lna_full_loglik_dev")

    def step_host(i):
        p_, i_, e_ = host[i % NBUF]
        _lib.check(L.lna_full_loglik(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), coal_h.data_ptr(),
                                     st_h.data_ptr(), None, None, None, None, None), "lna_full_loglik")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fp64_peak = L.lna_measure_fp64_peak(local_rank, 20000) if rank == 0 else 0.0

    for i in range(args.warmup):
        step_dev(i)
        step_host(i)
    torch.cuda.synchronize(dev)
    # sanity: device-resident and host-buffer paths agree
    step_dev(0)
    step_host(0)
    torch.cuda.synchronize(dev)
    assert torch.equal(coal_d.cpu(), coal_h), "device-resident and host-buffer results differ"
    n_sentinel = int((coal_h == -10000000.0).sum())

    # clocks: sample from a ~0.6 s untimed pre-load of the same kernel through the timed region
    clocks = Clocks(local_rank)
    clocks.start()
    t_load = time.perf_counter()
    i = 0
    while time.perf_counter() - t_load < 0.6:
        step_dev(i)
        i += 1
        if i % 16 == 0:
            torch.cuda.synchronize(dev)
    launches0 = pb.launch_count()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        evs[i][0].record(stream)
        step_dev(i)
        evs[i][1].record(stream)
    barrier()
    clk = clocks.stop()
    launches = pb.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- supplemental: the same launches with TWO batches in flight (two problem handles, two streams) ----
    # The headline batch is 2048 warp-loads on 592 SM sub-partitions (3.46 each): alone, a launch lasts as long as the
    # sub-partitions that got 4.  With the next batch's launch overlapping the tail the quantisation loss disappears; the
    # inputs rotate over NBUF = 4 distinct 65 MB batches (260 MB > L2), results go to two output buffers.
    pipelined = None
    try:
        pb2 = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init, device=local_rank)
        s2 = torch.cuda.Stream(dev)
        _lib.check(L.lna_problem_set_stream(pb2.h, C.c_void_p(s2.cuda_stream)), "set_stream")
        coal_d2 = torch.empty(B, dtype=torch.float64, device=dev)
        st_d2 = torch.empty(B, dtype=torch.int32, device=dev)

        def step_two(i):
            p_, i_, e_ = bufs[i % NBUF]
            h, cd, sd = (pb.h, coal_d, st_d) if i % 2 == 0 else (pb2.h, coal_d2, st_d2)
            _lib.check(L.lna_full_loglik_dev(h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cd.data_ptr(), sd.data_ptr(),
                                             None, None, None, None, None), "lna_full_loglik_dev")

        nst = max(args.steps, 8) * 2
        for i in range(4):
            step_two(i)
        barrier()
        torch.cuda.synchronize(dev)
        f0, f1, fj = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event())
        f0.record(stream)
        s2.wait_event(f0)
        for i in range(nst):
            step_two(i)
        fj.record(s2)
        stream.wait_event(fj)
        f1.record(stream)
        torch.cuda.synchronize(dev)
        tp = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pv = world * B * nst / (float(tp.item()) * 1e-3)
        assert torch.equal(coal_d2.cpu(), coal_d2.cpu()) and bool(torch.isfinite(coal_d2).all())
        pipelined = {"value": pv, "unit": UNIT, "launches": nst, "ms_per_launch": float(tp.item()) / nst,
                     "how": "two problem handles on two streams, consecutive launches overlap; 4 rotating 65 MB input batches"}
    except Exception as e:  # noqa: BLE001
        pipelined = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    # Every step copies its 65 MB of inputs from pinned host memory and reads its B likelihoods + statuses back.  Steps are
    # issued with the asynchronous form of the call (lna_full_loglik_async: same work, the H2D copies of step i+1 overlap
    # the last kernel of step i) and the region ends when the last result has landed in host memory (lna_problem_sync).
    coal_h2 = torch.empty(B, dtype=torch.float64).pin_memory()
    st_h2 = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_host_async(i):
        p_, i_, e_ = host[i % NBUF]
        cl, st_ = (coal_h, st_h) if i % 2 == 0 else (coal_h2, st_h2)
        _lib.check(L.lna_full_loglik_async(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cl.data_ptr(), st_.data_ptr()),
                   "lna_full_loglik_async")

    for i in range(2):
        step_host_async(i)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    step_host(0)
    ref0 = coal_h.clone()
    step_host_async(0)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    assert torch.equal(ref0, coal_h), "asynchronous and synchronous host-buffer calls differ"
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host_async(i)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i)
    torch.cuda.synchronize(dev)
    e2e_sync_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / float(t.item())
    e2e_sync_val = world * B * args.steps / e2e_sync_s
    h2d = sum(x.numel() * x.element_size() for x in host[0])
    d2h = coal_h.numel() * 8 + st_h.numel() * 4

    # ---- ESlice leg (batched General_MCMC_with_ESlice iterations), when the chain driver is built ----
    eslice = None
    if not args.no_eslice:
        try:
            eslice = bench_eslice(pb, g, args.chains, args.eslice_iters, rank, world, dev)
        except Exception as e:  # noqa: BLE001
            eslice = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- configs[0]: the constant_sim vignette's General_MCMC2 iteration body on resident chains (device-drawn streams) ----
    mcmc2 = None
    if not args.no_eslice:
        try:
            from conftest import Golden
            from lnaphylodyn_b200.chains import Chains, make_mcmc2_opts
            gc = Golden("constant_sim")
            pbc = api.Problem(gc.times, gc.x_r, gc.x_i, gc.gridsize, gc.t_correct, gc.init, device=local_rank)
            par0 = np.concatenate([[gc.x_r[0], float(gc.setting("control.alpha")[0]) * gc.x_r[0]], gc.setting("control.R0"),
                                   gc.setting("control.gamma"), gc.setting("control.ch"), gc.setting("control.hyper")])
            pr = [gc.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
            po = [gc.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
            o2 = make_mcmc2_opts(pr, po)
            mcmc2 = {"unit": "chain-iterations/s", "workload": "constant_sim vignette (E = 1057), General_MCMC2 body: 3 Metropolis "
                     "entries + hyper step + change-point slice update + trajectory slice update", "by_chains": {}}
            for Bc in (1, 1024, 4096):
                chc = Chains(pbc, np.tile(par0, (Bc, 1)), gc.setting("control.traj"))
                for it in range(6):
                    _lib.check(L.lna_chains_step_mcmc2_rng(chc.h, C.byref(o2), 7, it, rank * Bc), "mcmc2")
                pbc.sync()
                t0 = time.perf_counter()
                for it in range(6, 6 + args.eslice_iters):
                    _lib.check(L.lna_chains_step_mcmc2_rng(chc.h, C.byref(o2), 7, it, rank * Bc), "mcmc2")
                pbc.sync()
                dt = (time.perf_counter() - t0) / args.eslice_iters
                mcmc2["by_chains"][str(Bc)] = {"ms_per_iter": 1e3 * dt, "value": world * Bc / dt}
        except Exception as e:  # noqa: BLE001
            mcmc2 = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- configs[3]: SEIR (p = 3) parameter sweep, 16384 i.i.d. draws per sweep, FULL evaluations only ----
    seir = None
    if not args.no_seir:
        try:
            rng = np.random.default_rng(77 + rank)
            Bs = 16384
            prm = np.empty((Bs, 43))
            prm[:, 0] = np.exp(0.8 + 0.15 * rng.standard_normal(Bs))
            prm[:, 1] = np.exp(-0.7 + 0.1 * rng.standard_normal(Bs))
            prm[:, 2] = np.exp(-1.6 + 0.1 * rng.standard_normal(Bs))
            prm[:, 3:42] = np.exp(rng.standard_normal((Bs, 39)) / 30.0)
            prm[:, 42] = np.exp(3 + 0.2 * rng.standard_normal(Bs))
            pbs = api.Problem(g.times, g.x_r, [39, 3, 0, 2], g.gridsize, g.t_correct, g.init, model="SEIR", device=local_rank)
            _lib.check(L.lna_problem_set_stream(pbs.h, C.c_void_p(stream.cuda_stream)), "set_stream")
            d_prm = torch.from_numpy(prm).to(dev)
            d_ini = torch.tensor([[1e6, 4.0, 4.0]], dtype=torch.float64, device=dev).repeat(Bs, 1).contiguous()
            d_eps = (0.3 * torch.randn((Bs, 3 * 40), dtype=torch.float64, device=dev)).contiguous()
            d_out = torch.empty(Bs, dtype=torch.float64, device=dev)
            d_st = torch.empty(Bs, dtype=torch.int32, device=dev)
            run = lambda: _lib.check(L.lna_full_loglik_dev(pbs.h, Bs, d_prm.data_ptr(), d_ini.data_ptr(), d_eps.data_ptr(),
                                                           d_out.data_ptr(), d_st.data_ptr(), None, None, None, None, None), "seir")
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                run()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms_s = e0.elapsed_time(e1) / 10
            Wseir = 277.0 * (len(g.times) - 1) + 45.0 * 40 + 10.0 * 36
            seir = {"value": Bs / (ms_s * 1e-3), "unit": UNIT, "draws_per_sweep": Bs, "ms_per_sweep": ms_s,
                    "flops_per_eval": Wseir, "achieved_tflops": Bs / (ms_s * 1e-3) * Wseir / 1e12,
                    "sentinel_frac": float((d_out == -10000000.0).double().mean())}
        except Exception as e:  # noqa: BLE001
            seir = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        W = algorithmic_flops(g)
        achieved = (value / world) * W / 1e12
        peak = fp64_peak / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "chains_per_gpu": args.chains, "proposals_per_chain": args.proposals, "evals_per_step_per_gpu": B,
                       "l2": "256 MB flush write between timed iterations", "sentinel_evals_in_batch": n_sentinel},
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "lna_full_loglik_async per step + lna_problem_sync at the end (pinned host buffers)",
                    "value_synchronous_calls": e2e_sync_val},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak > 0 else None,
                         "traffic": NCU_TRAFFIC_BYTES_65536 if B == 65536 else None,
                         "traffic_unit": "bytes per launch (algorithmic: %d)" % ((h2d + d2h)),
                         "flops_per_eval": W,
                         "peak_source": "measured here: independent-DFMA microbenchmark (lna_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "hbm_algorithmic_GBps": (value / world) * (h2d + d2h) / B / 1e9},
        }
        if pipelined is not None:
            if "value" in pipelined and peak > 0:
                pipelined["frac"] = (pipelined["value"] / world) * W / 1e12 / peak
            line["two_batches_in_flight"] = pipelined
        if eslice is not None:
            line["eslice"] = eslice
        if mcmc2 is not None:
            line["mcmc2"] = mcmc2
        if seir is not None:
            line["seir_sweep"] = seir
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            impl, kind, what = cpu_impl()
            per = args.cpu_evals_per_core if args.cpu_evals_per_core > 0 else (512 if kind == "reference" else 4096)
            time_oracle(g, 8, cores, impl=impl)
            rate1, dt1 = time_oracle(g, max(64, per // 2), 1, impl=impl)
            rateN, dtN = time_oracle(g, per, cores, impl=impl)
            line["cpu_baseline"] = {"value": rateN, "unit": UNIT, "cores": cores, "kind": kind, "value_1core": rate1,
                                    "sample": f"{per} FULL evals/core x {cores} threads ({dtN:.1f}s) + "
                                              f"{max(64, per // 2)} on 1 core ({dt1:.1f}s); {what}"}
            if kind == "reference":  # the allocation-free C restatement beside it, for scale
                from oracle import oracle as orc
                rp1, _ = time_oracle(g, 1024, 1, impl=orc)
                rpN, _ = time_oracle(g, 2048, cores, impl=orc)
                line["cpu_baseline"]["port_value"] = rpN
                line["cpu_baseline"]["port_value_1core"] = rp1
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
