#!/usr/bin/env python
"""bench.py -- LNA+coalescent log-likelihood evaluations/sec (and ESlice chain-iterations/sec).

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1) prints ONE JSON
line on rank 0.  Workload = BASELINE.json configs[1] (Changpoint_sim vignette: SIR LNA, N=1e6,
2000 fine steps, gridsize 50, 39 R0 change points, the vignette's cached 342-tip genealogy E=376),
4096 chains.  One "step" of the likelihood leg = one batch of B FULL evaluations
(param, initial, OriginTraj) -> coalLog per GPU, B = 4096 chains x 16 slice proposals by default.

  value        : FULL evaluations/s, inputs already resident in HBM, CUDA-event timed on the launch stream
  e2e          : the same workload through the host-buffer C ABI call (pinned host inputs, H2D + kernel + D2H every step);
                 the call is the chain-shared form lna_full_loglik_shared_async: the 16 proposals of a chain share the chain's
                 initial state and latent increments, so a step ships 377 B per evaluation instead of 992
  roofline     : algorithmic FP64 flops (113/fine step + 21/interval + 10/cell) over the measured
                 independent-DFMA peak of this GPU, plus pipe_util = executed FP64 warp-instructions x 2 cycles over elapsed
                 cycles x 592 sub-partition pipes, from the instruction counts in profiles/r2_sass_fp64_counts.json
  eslice       : chain-iterations/s of the batched General_MCMC_with_ESlice iteration, with its own roofline object
                 (consumed evaluations x W over the peak; executed / consumed evaluations) and by_chains {4096, 16384, 65536}
  mcmc2        : configs[0] General_MCMC2 iteration on resident chains
  seir_sweep, sirs_sweep : configs[3], 16 384 parameter draws per likelihood sweep (SEIR via model="SEIR"; seasonal SIRS fused)
  ebola        : configs[2], the Ebola vignette's iteration as written, 8192 chains over the N GPUs (strong scaling)
  synthetic_sweep : configs[4], 5000-tip synthetic genealogy, chain counts 1k..64k per GPU
  summary_allgather : the one exchange of the path, ncclAllGather behind the C ABI (lna_comm_*), N > 1 only
  cpu_baseline : the reference's CPU implementation on the host cores, bounded sample: the reference's own
                 sources compiled verbatim on stand-in Armadillo/Rcpp headers (oracle/_ref, kind "reference") when
                 that library is present, else the C restatement (kind "port")

`--impl reference` times only that CPU arm (R/Rcpp/Armadillo are absent, so the reference's package build cannot
run; its C++ sources compile as they are against oracle/shim/).  `--workload ebola|synthetic` moves the headline legs to
configs[2] / configs[4]; `--chains` sets the chains per GPU.
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
WORKLOADS = {
    "changepoint": ("configs[1] Changpoint_sim vignette: SIR LNA N=1e6, 2000 fine steps, gridsize 50, 39 R0 change points, "
                    "cached 342-tip genealogy (E=376); FULL (param, initial, OriginTraj) -> coalLog evaluations"),
    "ebola": ("configs[2] Ebola Sierra Leone 2014 tree (bundled RData, 1010 tips, E=1774), change-point SIR LNA N=7e6, 2000 fine "
              "steps on [0, 1.36], gridsize 50, 39 change points; FULL evaluations"),
    "synthetic": ("configs[4] synthetic 5000-tip volz_thin_sir genealogy (E=5035) on the deterministic SIR trajectory N=1e6, "
                  "R0=2.2, gamma=0.2; 2000 fine steps, gridsize 50, 39 change points; FULL evaluations"),
}
# DRAM bytes per launch of the hot kernel at the default batch: dram__bytes_read.sum + dram__bytes_write.sum of ONE
# `ncu --set full` capture (a constant from that capture, not measured by this run)
NCU_TRAFFIC = {"bytes": 65071360 + 579840, "source": "dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture "
               "profiles/r2b_ncu_kernels.md (k_full_loglik_bal<false>, B = 65536, 148 x 512 threads; tools/headline_kernel.py): ncu "
               "cannot run inside the timed bench, so the figure is the capture's, not this run's"}
SMSP = 4  # FP64 pipes (SM sub-partitions) per SM
P3_NOTE = ("`achieved` uses SURVEY.md 8d's dense small-matrix ESTIMATE for p = 3 (~277 flops per fine step); the evidence is "
           "`achieved_executed` (flops of the instructions the kernel executes, counted in the SASS) and `pipe_util`")


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    """times, x_r, x_i, gridsize, t_correct, init + the chain state the legs start from and the vignette's options."""

    def __init__(self, name):
        from conftest import Golden, GOLDEN
        self.name = name
        self.prior_names = ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")
        self.prop_names = ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")
        if name == "changepoint":
            g = Golden("changepoint_sim")
            self.g = g
            self.times, self.x_r, self.x_i, self.gridsize, self.t_correct, self.init = (g.times, g.x_r, g.x_i, g.gridsize,
                                                                                         g.t_correct, g.init)
            self.par_center = g.final["par"].ravel()
            self.eps_center = g.final["OriginTraj"]
            self.par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                                        g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
            self.eps0 = g.setting("control.traj")
            self.prior = [g.setting("prior." + k) for k in self.prior_names]
            self.proposal = [g.setting("proposal." + k) for k in self.prop_names]
            self.entries = None
            return
        from lnaphylodyn_b200 import genealogy as gen
        if name == "ebola":
            z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
            s = gen.summarize_phylo(z["edge"], z["edge_length"], int(z["ntip"]))
            self.times, self.t_correct, N = np.linspace(0, 1.36, 2001), 1.36, 7e6
            self.par0 = np.concatenate([[N, 1.0, 2.0, 30.0], np.full(39, 0.975), [20.0]])
            self.prior = [[1, 0.5, 1, 1], [0.7, 0.5], [3, 0.1], [3, 0.15], [3, 0.2]]
            self.proposal = [[0.2], [0.05], [0.09], [0.1], [0.2]]
            self.entries = ((1, 1, 1), (3, 1, 3), (39 + 5, 0, 5))  # the vignette's parIdlist / isLoglist / priorIdlist as written
        else:
            s = np.load(os.path.join(GOLDEN, "synthetic_5000.npz"))
            self.times, self.t_correct, N = np.linspace(0, 100, 2001), 90.0, 1e6
            self.par0 = np.concatenate([[N, 1.0, 2.2, 0.2], np.ones(39), [20.0]])
            gq = Golden("changepoint_sim")  # the Changpoint_sim vignette's priors and proposals
            self.prior = [gq.setting("prior." + k) for k in self.prior_names]
            self.proposal = [gq.setting("proposal." + k) for k in self.prop_names]
            self.entries = None
        self.gridsize = 50
        self.init = gen.coal_lik_init(s["samp_times"], s["n_sampled"], s["coal_times"], self.times[::50])
        cht = self.times[::50][1:-1]
        self.x_r, self.x_i = np.concatenate([[N], cht]), [len(cht), 2, 0, 1]
        self.par_center, self.eps_center, self.eps0 = self.par0, np.zeros((40, 2)), np.zeros((40, 2))

    def problem(self, api, device):
        return api.Problem(self.times, self.x_r, self.x_i, self.gridsize, self.t_correct, self.init, device=device)

    def opts(self, update_traj=True):
        from lnaphylodyn_b200 import chains as chm
        ess = [0, 1, 0, 0, 1 if update_traj else 0, 1, 0]  # ESS_temp for i < burn1 (LNA_chpt_slice.R:773-774)
        if self.entries:
            return chm.make_opts_general(self.prior, self.proposal, ess, self.entries, hyper=False, update_traj=update_traj)
        return chm.make_opts(self.prior, self.proposal, ess, update_traj=update_traj, hyper_noop_mh=True)

    def flops_per_eval(self):
        """SURVEY.md section 8d, p = 2: 113 per fine step + 21 per LNA interval + 10 per coalescent cell (the 4-per-event
        term is hoisted into the per-genealogy precompute and not counted)."""
        n = len(self.times) - 1
        return 113.0 * n + 21.0 * (n // self.gridsize) + 10.0 * (int(np.asarray(self.init["ng"]).ravel()[0]) - 1)


def synth_inputs(w, n_chains, proposals, seed):
    """Synthetic slice proposals of the workload's shape: per chain an initial state and latent increments around the
    centre state, per proposal a parameter vector (R0, gamma and the change ratios perturbed)."""
    rng = np.random.default_rng(seed)
    c = w.par_center
    nch = len(c) - 5
    param = np.tile(c[2:], (n_chains, proposals, 1))
    param[:, :, 0] *= np.exp(0.05 * rng.standard_normal((n_chains, proposals)))
    param[:, :, 1] *= np.exp(0.05 * rng.standard_normal((n_chains, proposals)))
    param[:, :, 2:2 + nch] *= np.exp(0.02 * rng.standard_normal((n_chains, proposals, nch)))
    initial = np.tile(c[:2], (n_chains, 1))
    th = rng.uniform(-0.3, 0.3, size=(n_chains, 1, 1))
    eps = w.eps_center[None] * np.cos(th) + rng.standard_normal((n_chains, 40, 2)) * np.sin(th)
    return param, initial, eps


def sass_counts():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_sass_fp64_counts.json")))
    except Exception:  # noqa: BLE001
        return {}


def pipe_util(fp64_per_step, n_steps, evals, seconds, sm_count, sm_mhz, extra_fp64_per_eval=0.0):
    """Executed FP64 warp-instructions x 2 cycles over (elapsed cycles x SM sub-partition pipes).  A warp carries 32
    evaluations; `extra` covers the interval epilogues when known (else the time loop alone: a lower bound)."""
    if not (fp64_per_step and sm_mhz and seconds > 0):
        return None
    warp_instr = (evals / 32.0) * (fp64_per_step * n_steps + extra_fp64_per_eval)
    return warp_instr * 2.0 / (seconds * sm_mhz * 1e6 * sm_count * SMSP)


# ------------------------------------------------------------------------------------------------
# CPU arm (cpu_baseline leg and --impl reference)
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


def time_oracle(w, n_per_thread, threads, seed=123, impl=None):
    if impl is None:
        from oracle import oracle as impl
    orc = impl
    cfg, ini = orc.Cfg(w.x_r, w.x_i, len(w.par0) - 2), orc.Init(w.init)
    n = n_per_thread * threads
    param, initial, eps = synth_inputs(w, n, 1, seed)
    param, chunks = param[:, 0], []
    for i in range(threads):
        sl = slice(i * n_per_thread, (i + 1) * n_per_thread)
        chunks.append((param[sl], initial[sl], eps[sl]))
    outs = [None] * threads

    def work(i):
        outs[i] = orc.full_loglik_batch(cfg, ini, chunks[i][0], chunks[i][1], w.times, w.gridsize, chunks[i][2], w.t_correct)

    orc.lib()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return n / dt, dt


def bench_config(args, w):
    """The `config` object: identical in the b200 and the reference arm (the driver compares the two)."""
    return {"workload": WORKLOADS[w.name], "chains_per_gpu": args.chains, "proposals_per_chain": args.proposals,
            "evals_per_step_per_gpu": args.chains * args.proposals, "l2": "256 MB flush write between timed iterations"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = Workload(args.workload)
    cores = os.cpu_count() or 1
    impl, kind, what = cpu_impl()
    per = args.ref_evals_per_core if args.ref_evals_per_core > 0 else (96 if kind == "reference" else 1024)
    per = max(16, int(per))
    for _ in range(args.warmup):
        time_oracle(w, 16, cores, impl=impl)
    tot_t, tot_n = 0.0, 0
    for s in range(args.steps):
        rate, dt = time_oracle(w, per, cores, seed=1000 + s, impl=impl)
        tot_t += dt
        tot_n += per * cores
    val = tot_n / tot_t
    sample = (f"each step = {per} FULL evals per core x {cores} threads = {per * cores} of the step's "
              f"{args.chains * args.proposals} evaluations; {what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(args, w),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def cpu_reference_eslice(w, iters=12):
    """CPU leg for the ESlice metric: the restated reference driver (test infrastructure under oracle/, timed only -- never
    on the product path) on ONE host core, a bounded sample of chain-iterations.  Changpoint_sim workload only."""
    from lnaphylodyn_b200.chains import ITER_UNIF
    impl, kind, what = cpu_impl()
    if kind == "reference":
        driver = impl.driver()
    else:
        from oracle import driver
    g = w.g
    st = driver.init_state(g, w.par0, w.eps0)
    opts_b, opts = driver.opts_from_golden(g, update_traj=False), driver.opts_from_golden(g, update_traj=True)
    rng = np.random.default_rng(99)
    for _ in range(4):
        st = driver.eslice_iteration(st, g, opts_b, rng.standard_normal(123), rng.random(ITER_UNIF))
    t0 = time.perf_counter()
    for _ in range(iters):
        st = driver.eslice_iteration(st, g, opts, rng.standard_normal(123), rng.random(ITER_UNIF))
    dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": "chain-iterations/s", "cores": 1, "kind": kind,
            "sample": f"{iters} iterations of one chain, the R-level iteration body (oracle/driver.py) over: {what}"}


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
        except Exception:  # noqa: BLE001
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
        except Exception:  # noqa: BLE001
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
# legs
# ------------------------------------------------------------------------------------------------
class Ctx:
    """What every leg needs."""

    def __init__(self, args):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from lnaphylodyn_b200 import _lib, api
        self.C, self.torch, self.dist, self._lib, self.api = C, torch, dist, _lib, api
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        # one process per GPU: stay on the GPU's NUMA node before any pinned host buffer is allocated (e2e legs)
        self.host_binding = {"bound": False}
        if self.world > 1 and not os.environ.get("LNA_NO_NUMA_BIND"):
            from lnaphylodyn_b200 import comm as _comm
            pr = torch.cuda.get_device_properties(self.dev)
            self.host_binding = _comm.bind_host_to_device_numa(getattr(pr, "pci_domain_id", 0), getattr(pr, "pci_bus_id", -1),
                                                               getattr(pr, "pci_device_id", 0))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.L = _lib.lib()
        self.stream = torch.cuda.current_stream(self.dev)
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.sass = sass_counts()
        self.peak = 0.0
        self.sm_mhz = None

    def problem(self, w):
        pb = w.problem(self.api, self.local_rank)
        self._lib.check(self.L.lna_problem_set_stream(pb.h, self.C.c_void_p(self.stream.cuda_stream)), "set_stream")
        return pb

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)


def eslice_leg(cx, w, pb, B, iters, with_e2e=False, with_cpu=False):
    """ESlice chain-iterations/s: B chains per GPU from the workload's initial state, the iteration on device-drawn streams
    (lna_chains_step_rng: Philox in HBM, replayed from a CUDA graph), device-timed; optionally the host-stream e2e forms."""
    torch, L, C = cx.torch, cx.L, cx.C
    from lnaphylodyn_b200._lib import check
    from lnaphylodyn_b200.chains import Chains, ITER_UNIF, ITER_UNIF_GENERAL
    ch = Chains(pb, np.tile(w.par0, (B, 1)), w.eps0)
    ob, of = w.opts(update_traj=False), w.opts(update_traj=True)
    seed, off = 20261020, cx.rank * B
    for i in range(4):  # burn-in like the vignette's first stage, then warm-up of the full iteration (incl. graph capture)
        ch.step_rng(ob, seed, i, off)
    for i in range(4, 10):
        ch.step_rng(of, seed, i, off)
    ch.sync()
    before, (ex0, wx0) = ch.get(), ch.exec_counts()
    cx.barrier()
    l0 = pb.launch_count()
    e0, e1 = cx.events()
    e0.record(cx.stream)
    for i in range(iters):
        ch.step_rng(of, seed, 10 + i, off)
    ch.join()
    e1.record(cx.stream)
    torch.cuda.synchronize(cx.dev)
    ms = e0.elapsed_time(e1)
    launches = pb.launch_count() - l0
    after, (ex1, wx1) = ch.get(), ch.exec_counts()
    consumed = float((after["n_full_evals"] - before["n_full_evals"]).sum())
    cheap = float((after["n_cheap_evals"] - before["n_cheap_evals"]).sum())
    ms, = cx.max_over_ranks(ms)
    consumed, cheap, executed, exec_warps = cx.sum_over_ranks(consumed, cheap, ex1 - ex0, wx1 - wx0)
    tot = cx.world * B * iters
    W = w.flops_per_eval()
    out = {"value": tot / (ms * 1e-3), "unit": "chain-iterations/s", "chains_per_gpu": B, "iters": iters,
           "ms_per_iter": ms / iters, "stream_groups": ch.groups, "streams": "device-drawn (Philox4x32-10, lna_chains_step_rng)",
           "mean_full_evals_per_iter": consumed / tot, "mean_cheap_evals_per_iter": cheap / tot,
           "consumed_full_evals_per_s": consumed / (ms * 1e-3), "executed_full_evals_per_iter": executed / tot,
           "gpu_launches": int(launches), "finite_coalLog": bool(np.all(np.isfinite(after["coalLog"])))}
    if cx.peak > 0:
        ach = consumed / cx.world / (ms * 1e-3) * W / 1e12
        cnt = cx.sass.get("cand_ess_par_w8", {})
        out["roofline"] = {"bound": "fp64", "achieved": ach, "peak": cx.peak, "unit": "TFLOP/s", "frac": ach / cx.peak,
                           "what": "CONSUMED FULL evaluations (what the sequential reference loop evaluates) x flops_per_eval "
                                   "/ time; speculative evaluations are not counted as work",
                           "flops_per_eval": W, "executed_over_consumed": executed / max(consumed, 1.0),
                           "pipe_util": pipe_util(cnt.get("fp64_per_step"), len(w.times) - 1, 32.0 * exec_warps / cx.world,
                                                  ms * 1e-3, cx.sm_count, cx.sm_mhz),
                           "pipe_util_what": "executed warp-evaluations (idle lanes cost the pipe the same) x FP64 "
                                             "instructions of the time loop x 2 cycles / (elapsed cycles x 592 pipes)"}
    if with_e2e:
        general = of.n_mh > 0
        nu = ITER_UNIF_GENERAL if general else ITER_UNIF
        nrm_h = torch.randn((B, ch.n_norm), dtype=torch.float64).pin_memory()
        unf_h = torch.rand((B, nu), dtype=torch.float64).clamp_(1e-300, 1 - 1e-16).pin_memory()
        cl_h = torch.empty(B, dtype=torch.float64).pin_memory()
        lo_h = torch.empty(B, dtype=torch.float64).pin_memory()
        for i in range(3):  # first-touch of the pinned buffers and of the host-pointer staging path, untimed
            check(L.lna_chains_step(ch.h, C.byref(of), nrm_h.data_ptr(), unf_h.data_ptr()), "lna_chains_step")
            check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
        t0 = time.perf_counter()
        for i in range(iters):
            check(L.lna_chains_step(ch.h, C.byref(of), nrm_h.data_ptr(), unf_h.data_ptr()), "lna_chains_step")
            check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
        e2e_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(iters):  # streams drawn on the device: only the summaries cross PCIe
            ch.step_rng(of, seed, 1000 + i, off)
            check(L.lna_chains_get(ch.h, None, None, None, cl_h.data_ptr(), lo_h.data_ptr(), None), "lna_chains_get")
        rng_s = time.perf_counter() - t0
        e2e_s, rng_s = cx.max_over_ranks(e2e_s, rng_s)
        out.update({"e2e_value": tot / e2e_s, "e2e_device_rng_value": tot / rng_s,
                    "h2d_bytes_per_iter": int(nrm_h.numel() * 8 + unf_h.numel() * 8), "d2h_bytes_per_iter": int(B * 16)})
    if with_cpu and cx.rank == 0 and w.name == "changepoint":
        out["cpu_baseline"] = cpu_reference_eslice(w)
    del ch
    return out


def full_leg_quick(cx, w, pb, n_chains, proposals, reps=10):
    """FULL evaluations/s of one more workload / batch size, device-resident chain-shared inputs (no flush: inputs > L2 or
    compute-bound either way)."""
    torch, L = cx.torch, cx.L
    param, initial, eps = synth_inputs(w, n_chains, proposals, 5 + cx.rank)
    B = n_chains * proposals
    p_ = torch.from_numpy(param.reshape(B, -1)).to(cx.dev)
    i_ = torch.from_numpy(initial).to(cx.dev)
    e_ = torch.from_numpy(np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))).to(cx.dev)
    cl = torch.empty(B, dtype=torch.float64, device=cx.dev)
    st = torch.empty(B, dtype=torch.int32, device=cx.dev)
    run = lambda: cx._lib.check(L.lna_full_loglik_shared_dev(pb.h, n_chains, proposals, p_.data_ptr(), i_.data_ptr(),
                                                             e_.data_ptr(), cl.data_ptr(), st.data_ptr()), "shared_dev")
    for _ in range(3):
        run()
    e0, e1 = cx.events()
    cx.barrier()
    e0.record(cx.stream)
    for _ in range(reps):
        run()
    e1.record(cx.stream)
    torch.cuda.synchronize(cx.dev)
    ms, = cx.max_over_ranks(e0.elapsed_time(e1) / reps)
    val = cx.world * B / (ms * 1e-3)
    out = {"value": val, "unit": UNIT, "evals_per_launch_per_gpu": B, "us_per_launch": ms * 1e3,
           "sentinel_frac": float((cl == -10000000.0).double().mean())}
    if cx.peak > 0:
        out["roofline_frac"] = val / cx.world * w.flops_per_eval() / 1e12 / cx.peak
    return out


def p3_legs(cx, g):
    """configs[3]: 16 384 i.i.d. parameter draws per likelihood sweep, FULL evaluations only (no sampler), p = 3."""
    torch, L, api, _lib = cx.torch, cx.L, cx.api, cx._lib
    out = {}
    Bs, n, k = 16384, len(g.times) - 1, 40
    rng = np.random.default_rng(77 + cx.rank)

    def time_it(run, reps=10):
        for _ in range(3):
            run()
        e0, e1 = cx.events()
        e0.record(cx.stream)
        for _ in range(reps):
            run()
        e1.record(cx.stream)
        torch.cuda.synchronize(cx.dev)
        return e0.elapsed_time(e1) / reps

    def roof(key, ms, flops_alg, note):
        cnt = cx.sass.get(key, {})
        r = {"bound": "fp64", "unit": "TFLOP/s", "peak": cx.peak, "flops_per_eval_algorithmic": flops_alg,
             "flops_note": note}
        if cx.peak > 0:
            r["achieved"] = Bs / (ms * 1e-3) * flops_alg / 1e12
            r["frac"] = r["achieved"] / cx.peak
            if cnt:
                r["executed_flops_per_step"] = cnt["executed_flops_per_step"]
                r["achieved_executed"] = Bs / (ms * 1e-3) * cnt["executed_flops_per_step"] * n / 1e12
                r["frac_executed"] = r["achieved_executed"] / cx.peak
            r["pipe_util"] = pipe_util(cnt.get("fp64_per_step"), n, Bs, ms * 1e-3, cx.sm_count, cx.sm_mhz)
            r["regime"] = ("512 warps on 592 sub-partition pipes: one trajectory's latency bounds the launch, not the pipe "
                           "(pipe_util is the honest figure)")
        return r
    # ---- SEIR through the generic engine (model = "SEIR", LNA_functional.cpp:156-198, 276-344) ----
    try:
        prm = np.empty((Bs, 43))
        prm[:, 0] = np.exp(0.8 + 0.15 * rng.standard_normal(Bs))
        prm[:, 1] = np.exp(-0.7 + 0.1 * rng.standard_normal(Bs))
        prm[:, 2] = np.exp(-1.6 + 0.1 * rng.standard_normal(Bs))
        prm[:, 3:42] = np.exp(rng.standard_normal((Bs, 39)) / 30.0)
        prm[:, 42] = np.exp(3 + 0.2 * rng.standard_normal(Bs))
        pbs = api.Problem(g.times, g.x_r, [39, 3, 0, 2], g.gridsize, g.t_correct, g.init, model="SEIR", device=cx.local_rank)
        _lib.check(L.lna_problem_set_stream(pbs.h, cx.C.c_void_p(cx.stream.cuda_stream)), "set_stream")
        d_prm = torch.from_numpy(prm).to(cx.dev)
        d_ini = torch.tensor([[1e6, 4.0, 4.0]], dtype=torch.float64, device=cx.dev).repeat(Bs, 1).contiguous()
        d_eps = (0.3 * torch.randn((Bs, 3 * 40), dtype=torch.float64, device=cx.dev)).contiguous()
        d_out = torch.empty(Bs, dtype=torch.float64, device=cx.dev)
        d_st = torch.empty(Bs, dtype=torch.int32, device=cx.dev)
        ms = time_it(lambda: _lib.check(L.lna_full_loglik_dev(pbs.h, Bs, d_prm.data_ptr(), d_ini.data_ptr(), d_eps.data_ptr(),
                                                              d_out.data_ptr(), d_st.data_ptr(), None, None, None, None, None), "seir"))
        Wseir = 277.0 * n + 45.0 * k + 10.0 * 36
        out["seir_sweep"] = {"value": Bs / (ms * 1e-3), "unit": UNIT, "draws_per_sweep": Bs, "ms_per_sweep": ms,
                             "sentinel_frac": float((d_out == -10000000.0).double().mean()),
                             "roofline": roof("seir_full_nostore", ms, Wseir, P3_NOTE)}
    except Exception as e:  # noqa: BLE001
        out["seir_sweep"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    # ---- seasonal SIRS, fused ODE2 -> SIRS_KOM_Filter -> log_like_trajSIRS (src/SIRS_period.cpp) ----
    try:
        t = np.linspace(0, 100, 2001)
        pbq = api.Problem(t, [40.0], [0, 4, 0, 1], 50, 90.0, None, model="SIRS_period", device=cx.local_rank)
        _lib.check(L.lna_problem_set_stream(pbq.h, cx.C.c_void_p(cx.stream.cuda_stream)), "set_stream")
        prm = np.column_stack([3e-6 * np.exp(0.1 * rng.standard_normal(Bs)), 0.2 * np.exp(0.1 * rng.standard_normal(Bs)),
                               0.02 * np.exp(0.1 * rng.standard_normal(Bs)), rng.uniform(0.1, 0.5, Bs)])
        init = np.array([9e4, 100.0, 9900.0])
        # the "data": the ODE at the true parameters (product API) + a perturbation on the conserved-population plane
        truth = api.SIRS_loglik_sweep(np.array([[3e-6, 0.2, 0.02, 0.3]]), init, t, None, 50, 90.0, 40.0, want=("Ode",),
                                      device=cx.local_rank)["Ode"][0]
        dev_ = np.random.default_rng(5).standard_normal((41, 3)) * 30.0
        dev_[:, 2] = -dev_[:, 0] - dev_[:, 1]
        dev_[0] = 0
        sde = truth.copy()
        sde[:, 1:] += dev_
        d_prm = torch.from_numpy(prm).to(cx.dev)
        d_ini = torch.from_numpy(init).to(cx.dev)
        d_sde = torch.from_numpy(np.ascontiguousarray(sde.T)).to(cx.dev)  # nrow x 4 column-major
        d_out = torch.empty(Bs, dtype=torch.float64, device=cx.dev)
        d_st = torch.empty(Bs, dtype=torch.int32, device=cx.dev)
        ms = time_it(lambda: _lib.check(L.lna_sirs_loglik_dev(pbq.h, Bs, d_prm.data_ptr(), d_ini.data_ptr(), 1, d_sde.data_ptr(), 0,
                                                              d_out.data_ptr(), d_st.data_ptr(), None, None, None), "sirs"))
        Wsirs = 277.0 * n + 60.0 * k
        out["sirs_sweep"] = {"value": Bs / (ms * 1e-3), "unit": UNIT, "draws_per_sweep": Bs, "ms_per_sweep": ms,
                             "finite_frac": float(torch.isfinite(d_out).double().mean()),
                             "roofline": roof("sirs_full_nostore", ms, Wsirs, P3_NOTE)}
    except Exception as e:  # noqa: BLE001
        out["sirs_sweep"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="changepoint", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=4096, help="concurrent chains per GPU")
    ap.add_argument("--proposals", type=int, default=16, help="slice proposals per chain in one likelihood batch")
    ap.add_argument("--ref-evals-per-core", type=int, default=0, help="0 = 96 (reference as written) / 1024 (port)")
    ap.add_argument("--cpu-evals-per-core", type=int, default=0, help="0 = 512 (reference as written) / 4096 (port)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eslice", action="store_true")
    ap.add_argument("--no-p3", "--no-seir", dest="no_p3", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[4] / by_chains legs")
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

    cx = Ctx(args)
    torch, L, C, _lib, api = cx.torch, cx.L, cx.C, cx._lib, cx.api
    rank, world, dev, stream = cx.rank, cx.world, cx.dev, cx.stream
    w = Workload(args.workload)
    nC, m = args.chains, args.proposals
    B = nC * m
    pb = cx.problem(w)

    # ---- inputs: NBUF distinct batches (this rank's shard of the global batch), per-item and chain-shared layouts ----
    NBUF = 4
    bufs, host_sh, dev_sh = [], [], []
    for i in range(NBUF):
        param, initial, eps = synth_inputs(w, nC, m, seed=10_000 * (rank + 1) + i)
        eps_t = np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))  # item = k x p column-major
        ph = torch.from_numpy(param.reshape(B, -1)).pin_memory()
        ih, eh = torch.from_numpy(initial).pin_memory(), torch.from_numpy(eps_t.reshape(nC, -1)).pin_memory()
        host_sh.append((ph, ih, eh))
        pd, idv, ed = ph.to(dev), ih.to(dev), eh.to(dev)
        dev_sh.append((pd, idv, ed))
        # the per-item call's layout: every evaluation carries its own copy of the chain's initial state / increments
        bufs.append((pd, idv.repeat_interleave(m, dim=0).contiguous(), ed.repeat_interleave(m, dim=0).contiguous()))
    coal_d = torch.empty(B, dtype=torch.float64, device=dev)
    st_d = torch.empty(B, dtype=torch.int32, device=dev)
    coal_s = torch.empty(B, dtype=torch.float64, device=dev)
    st_s = torch.empty(B, dtype=torch.int32, device=dev)
    coal_h = [torch.empty(B, dtype=torch.float64).pin_memory() for _ in range(2)]
    st_h = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def step_dev(i, cl=coal_d, st=st_d, h=None):
        p_, i_, e_ = bufs[i % NBUF]
        _lib.check(L.lna_full_loglik_dev(h or pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cl.data_ptr(),
                                         st.data_ptr(), None, None, None, None, None), "lna_full_loglik_dev")

    def step_shared_dev(i):
        p_, i_, e_ = dev_sh[i % NBUF]
        _lib.check(L.lna_full_loglik_shared_dev(pb.h, nC, m, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), coal_s.data_ptr(),
                                                st_s.data_ptr()), "lna_full_loglik_shared_dev")

    def step_host(i, asynchronous=True):
        p_, i_, e_ = host_sh[i % NBUF]
        fn = L.lna_full_loglik_shared_async if asynchronous else L.lna_full_loglik_shared
        _lib.check(fn(pb.h, nC, m, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), coal_h[i % 2].data_ptr(),
                      st_h[i % 2].data_ptr()), "lna_full_loglik_shared")

    cx.peak = (L.lna_measure_fp64_peak(cx.local_rank, 20000) / 1e12) if rank == 0 else 0.0
    if world > 1:
        cx.peak, = cx.max_over_ranks(cx.peak)

    for i in range(args.warmup):
        step_dev(i)
        step_shared_dev(i)
        step_host(i)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    torch.cuda.synchronize(dev)
    # sanity: per-item, chain-shared device-resident and chain-shared host-buffer calls agree bit for bit
    step_dev(0)
    step_shared_dev(0)
    step_host(0, asynchronous=False)
    step_host(1)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    torch.cuda.synchronize(dev)
    assert torch.equal(coal_d, coal_s), "per-item and chain-shared results differ"
    assert torch.equal(coal_d.cpu(), coal_h[0]), "device-resident and host-buffer results differ"
    step_dev(1)
    torch.cuda.synchronize(dev)
    assert torch.equal(coal_d.cpu(), coal_h[1]), "asynchronous host-buffer results differ"
    n_sentinel = int((coal_h[0] == -10000000.0).sum())

    # ---- headline: clocks sampled from a ~0.6 s untimed pre-load of the same kernel through the timed region ----
    clocks = Clocks(cx.local_rank)
    clocks.start()
    t_load = time.perf_counter()
    i = 0
    while time.perf_counter() - t_load < 0.6:
        step_dev(i)
        i += 1
        if i % 16 == 0:
            torch.cuda.synchronize(dev)
    launches0 = pb.launch_count()
    cx.barrier()
    evs = [cx.events() for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        evs[i][0].record(stream)
        step_dev(i)
        evs[i][1].record(stream)
    cx.barrier()
    clk = clocks.stop()
    cx.sm_mhz = clk.get("sm_mhz") or clk.get("sm_max_mhz")
    launches = pb.launch_count() - launches0
    ms_total, = cx.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    value = world * B * args.steps / (ms_total * 1e-3)
    # the same launches on the chain-shared layout (what e2e ships): same kernel, 2.6 x fewer input bytes
    evs = [cx.events() for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        evs[i][0].record(stream)
        step_shared_dev(i)
        evs[i][1].record(stream)
    cx.barrier()
    ms_shared, = cx.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    value_shared = world * B * args.steps / (ms_shared * 1e-3)

    # ---- supplemental: the same launches with TWO batches in flight (two problem handles, two streams) ----
    # The headline batch is 2048 warp-loads on 592 SM sub-partitions (3.46 each).  The plain kernel lasted as long as the
    # sub-partitions that got 4 (0.510 ms); the balanced kernel (k_full_loglik_bal: each SM's loads cut into four equal shares)
    # brings a lone launch to 3.5 (0.455 ms).  What a second batch in flight still adds: its launch fills the 13-load SMs' tails.
    pipelined = None
    try:
        pb2 = w.problem(api, cx.local_rank)
        s2 = torch.cuda.Stream(dev)
        _lib.check(L.lna_problem_set_stream(pb2.h, C.c_void_p(s2.cuda_stream)), "set_stream")
        coal_d2 = torch.empty(B, dtype=torch.float64, device=dev)
        st_d2 = torch.empty(B, dtype=torch.int32, device=dev)

        def step_two(i):
            if i % 2 == 0:
                step_dev(i)
            else:
                step_dev(i, coal_d2, st_d2, pb2.h)

        nst = max(args.steps, 8) * 2
        for i in range(4):
            step_two(i)
        cx.barrier()
        f0, f1, fj = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event())
        f0.record(stream)
        s2.wait_event(f0)
        for i in range(nst):
            step_two(i)
        fj.record(s2)
        stream.wait_event(fj)
        f1.record(stream)
        torch.cuda.synchronize(dev)
        tp, = cx.max_over_ranks(f0.elapsed_time(f1))
        # both handles' last results against a plain single launch on the same input batches
        ref_even, ref_odd = torch.empty_like(coal_d), torch.empty_like(coal_d)
        step_dev(nst - 2, ref_even, st_s)
        step_dev(nst - 1, ref_odd, st_s)
        torch.cuda.synchronize(dev)
        assert torch.equal(coal_d, ref_even) and torch.equal(coal_d2, ref_odd), "two-stream results differ from single launches"
        pipelined = {"value": world * B * nst / (tp * 1e-3), "unit": UNIT, "launches": nst, "ms_per_launch": tp / nst,
                     "how": "two problem handles on two streams, consecutive launches overlap; 4 rotating input batches; "
                            "results checked against single launches"}
    except Exception as e:  # noqa: BLE001
        pipelined = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    # Every step copies its inputs from pinned host memory (chain-shared layout) and reads its B likelihoods + statuses back.
    # Steps are issued with the asynchronous form of the call (the H2D copies of step i+1 overlap the kernel of step i) and
    # the region ends when the last result has landed in host memory (lna_problem_sync).
    n_e2e = max(args.steps, 50)  # (the region also holds one pipeline fill and drain: ~1 step)
    cx.barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        step_host(i)
    _lib.check(L.lna_problem_sync(pb.h), "sync")
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    cx.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i, asynchronous=False)
    torch.cuda.synchronize(dev)
    e2e_sync_s = time.perf_counter() - t0
    e2e_s, e2e_sync_s = cx.max_over_ranks(e2e_s, e2e_sync_s)
    e2e_val = world * B * n_e2e / e2e_s
    e2e_sync_val = world * B * args.steps / e2e_sync_s
    h2d = sum(x.numel() * x.element_size() for x in host_sh[0])
    d2h = coal_h[0].numel() * 8 + st_h[0].numel() * 4

    # ---- ESlice leg (batched General_MCMC_with_ESlice iterations) ----
    eslice = None
    if not args.no_eslice:
        try:
            eslice = eslice_leg(cx, w, pb, nC, args.eslice_iters, with_e2e=True, with_cpu=not args.no_cpu_baseline)
            if not args.no_extra:
                eslice["by_chains"] = {}
                for Bc in (4096, 16384, 65536):
                    r = eslice if Bc == nC else eslice_leg(cx, w, pb, Bc, max(6, args.eslice_iters // 2))
                    eslice["by_chains"][str(Bc)] = {k: r.get(k) for k in ("value", "ms_per_iter", "stream_groups",
                                                                           "mean_full_evals_per_iter")}
                    if "roofline" in r:
                        eslice["by_chains"][str(Bc)].update({"roofline_frac": r["roofline"]["frac"],
                                                             "executed_over_consumed": r["roofline"]["executed_over_consumed"],
                                                             "pipe_util": r["roofline"]["pipe_util"]})
        except Exception as e:  # noqa: BLE001
            eslice = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    # ---- configs[0]: the constant_sim vignette's General_MCMC2 iteration body on resident chains ----
    mcmc2 = None
    if not args.no_eslice:
        try:
            from conftest import Golden
            from lnaphylodyn_b200.chains import Chains, make_mcmc2_opts
            gc = Golden("constant_sim")
            pbc = api.Problem(gc.times, gc.x_r, gc.x_i, gc.gridsize, gc.t_correct, gc.init, device=cx.local_rank)
            par0 = np.concatenate([[gc.x_r[0], float(gc.setting("control.alpha")[0]) * gc.x_r[0]], gc.setting("control.R0"),
                                   gc.setting("control.gamma"), gc.setting("control.ch"), gc.setting("control.hyper")])
            pr = [gc.setting("prior." + k) for k in w.prior_names]
            po = [gc.setting("proposal." + k) for k in w.prop_names]
            o2 = make_mcmc2_opts(pr, po)
            Wc = 113.0 * 2000 + 21.0 * 40 + 10.0 * (int(gc.init["ng"][0]) - 1)
            mcmc2 = {"unit": "chain-iterations/s", "workload": "constant_sim vignette (E = 1057), General_MCMC2 body: 3 Metropolis "
                     "entries + hyper step + change-point slice update + trajectory slice update", "by_chains": {}}
            for Bc in (1, 1024, 4096):
                chc = Chains(pbc, np.tile(par0, (Bc, 1)), gc.setting("control.traj"))
                for it in range(6):
                    _lib.check(L.lna_chains_step_mcmc2_rng(chc.h, C.byref(o2), 7, it, rank * Bc), "mcmc2")
                chc.sync()
                n0 = chc.get()["n_full_evals"].sum()
                t0 = time.perf_counter()
                for it in range(6, 6 + args.eslice_iters):
                    _lib.check(L.lna_chains_step_mcmc2_rng(chc.h, C.byref(o2), 7, it, rank * Bc), "mcmc2")
                chc.sync()
                dt = (time.perf_counter() - t0) / args.eslice_iters
                cons = float(chc.get()["n_full_evals"].sum() - n0) / args.eslice_iters
                rec = {"ms_per_iter": 1e3 * dt, "value": world * Bc / dt, "consumed_full_evals_per_iter": cons / Bc}
                if cx.peak > 0:
                    rec["roofline_frac"] = cons / dt * Wc / 1e12 / cx.peak
                mcmc2["by_chains"][str(Bc)] = rec
                del chc
        except Exception as e:  # noqa: BLE001
            mcmc2 = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- configs[3]: p = 3 parameter sweeps ----
    p3 = {} if args.no_p3 else p3_legs(cx, Workload("changepoint") if w.name != "changepoint" else w)

    # ---- configs[2] and configs[4] on the default run (compact) ----
    extra = {}
    if not args.no_extra and not args.no_eslice and w.name == "changepoint":
        try:
            we = Workload("ebola")
            pbe = cx.problem(we)
            total = 8192
            per = max(total // world, 1)
            r = eslice_leg(cx, we, pbe, per, args.eslice_iters)
            r.update({"scaling": "strong", "chains_total": per * world,
                      "what": "the Ebola vignette's iteration AS WRITTEN (Update_Param_each_Norm entries S0 / R0 / no-op hyper as "
                              "a 7-candidate tree, parameter and trajectory slice updates), 8192 chains over the GPUs of the run",
                      "limiter": "per-GPU batches of <= 2048 chains leave the SM sub-partitions <= 1-2 warps: the iteration "
                                 "lasts ~3 single-evaluation latencies whatever the chain count, so strong scaling flattens "
                                 "at the latency floor"})
            r["full_evals"] = full_leg_quick(cx, we, pbe, per, m)
            extra["ebola"] = r
        except Exception as e:  # noqa: BLE001
            extra["ebola"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        try:
            ws = Workload("synthetic")
            pbs_ = cx.problem(ws)
            sw = {"workload": WORKLOADS["synthetic"], "by_chains": {}}
            for Bc in (1024, 4096, 16384, 65536):
                r = eslice_leg(cx, ws, pbs_, Bc, max(6, args.eslice_iters // 2))
                f = full_leg_quick(cx, ws, pbs_, Bc, 1)
                sw["by_chains"][str(Bc)] = {"eslice_chain_iters_per_s": r["value"], "eslice_ms_per_iter": r["ms_per_iter"],
                                            "eslice_roofline_frac": r.get("roofline", {}).get("frac"),
                                            "full_evals_per_s": f["value"], "full_us_per_batch": f["us_per_launch"],
                                            "full_roofline_frac": f.get("roofline_frac")}
            extra["synthetic_sweep"] = sw
        except Exception as e:  # noqa: BLE001
            extra["synthetic_sweep"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    # ---- the one exchange of the path: all-gather of the per-chain summary rows behind the C ABI ----
    gather = None
    if world > 1:
        try:
            from lnaphylodyn_b200 import comm
            gather = comm.bench_allgather(cx, pb, nC)
        except Exception as e:  # noqa: BLE001
            gather = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    host_binding = cx.host_binding
    if world > 1:  # every rank's NUMA node, for the record
        allb = [None] * world
        cx.dist.all_gather_object(allb, cx.host_binding)
        host_binding = {"what": "each rank bound to the CPUs of its GPU's NUMA node before allocating pinned host buffers "
                                "(lnaphylodyn_b200.comm.bind_host_to_device_numa)", "ranks": allb}
    if rank == 0:
        W = w.flops_per_eval()
        achieved = (value / world) * W / 1e12
        peak = cx.peak
        cnt = cx.sass.get("sir_full_balanced_nostore", cx.sass.get("sir_full_nostore", {}))
        n = len(w.times) - 1
        per_item_bytes = sum(x.numel() * x.element_size() for x in bufs[0]) + B * 12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args, w),
            "sentinel_evals_in_batch": n_sentinel, "clocks": clk,
            "value_chain_shared_inputs": value_shared,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                    "call": "lna_full_loglik_shared_async per step + lna_problem_sync at the end (pinned host buffers; the 16 "
                            "proposals of a chain share the chain's initial state and OriginTraj: 377 B per evaluation, the "
                            "per-item call ships 992)",
                    "frac_of_device_resident": e2e_val / value, "value_synchronous_calls": e2e_sync_val},
            "gpu_launches": int(launches), "host_binding": host_binding,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak > 0 else None,
                         "pipe_util": pipe_util(cnt.get("fp64_per_step"), n, B * args.steps, ms_total * 1e-3, cx.sm_count, cx.sm_mhz),
                         "pipe_util_what": "executed FP64 warp-instructions of the time loop (profiles/r2_sass_fp64_counts.json: "
                                           f"{cnt.get('fp64_per_step')} per fine step; epilogues not counted) x 2 cycles / "
                                           "(elapsed cycles at the sampled SM clock x 592 sub-partition pipes)",
                         "executed_flops_per_step": cnt.get("executed_flops_per_step"),
                         "achieved_executed": (value / world) * (cnt.get("executed_flops_per_step", 0) * n) / 1e12,
                         "traffic": NCU_TRAFFIC["bytes"] if (B == 65536 and w.name == "changepoint") else None,
                         "traffic_source": NCU_TRAFFIC["source"],
                         "traffic_algorithmic": per_item_bytes, "flops_per_eval": W,
                         "peak_source": "measured here: independent-DFMA microbenchmark (lna_measure_fp64_peak, 16 warps per "
                                        "sub-partition, constant-bank operands); MEASURED_PEAKS.json has no FP64 entry; "
                                        "nominal 148 SM x 64 lanes x 2 x SM clock",
                         "hbm_algorithmic_GBps": (value / world) * per_item_bytes / B / 1e9},
        }
        if pipelined is not None:
            if "value" in pipelined and peak > 0:
                pipelined["frac"] = (pipelined["value"] / world) * W / 1e12 / peak
                pipelined["pipe_util"] = pipe_util(cnt.get("fp64_per_step"), n, B, pipelined["ms_per_launch"] * 1e-3, cx.sm_count,
                                                   cx.sm_mhz)
            line["two_batches_in_flight"] = pipelined
        if eslice is not None:
            line["eslice"] = eslice
        if mcmc2 is not None:
            line["mcmc2"] = mcmc2
        line.update(p3)
        line.update(extra)
        if gather is not None:
            line["summary_allgather"] = gather
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            impl, kind, what = cpu_impl()
            per = args.cpu_evals_per_core if args.cpu_evals_per_core > 0 else (512 if kind == "reference" else 4096)
            time_oracle(w, 8, cores, impl=impl)
            rate1, dt1 = time_oracle(w, max(64, per // 2), 1, impl=impl)
            rateN, dtN = time_oracle(w, per, cores, impl=impl)
            line["cpu_baseline"] = {"value": rateN, "unit": UNIT, "cores": cores, "kind": kind, "value_1core": rate1,
                                    "sample": f"{per} FULL evals/core x {cores} threads ({dtN:.1f}s) + "
                                              f"{max(64, per // 2)} on 1 core ({dt1:.1f}s); {what}"}
            if kind == "reference":  # the allocation-free C restatement beside it: always quote both ratios
                from oracle import oracle as orc
                rp1, _ = time_oracle(w, 1024, 1, impl=orc)
                rpN, _ = time_oracle(w, 2048, cores, impl=orc)
                line["cpu_baseline"].update({"port_value": rpN, "port_value_1core": rp1,
                                             "gpu_over_reference_as_written": value / rateN, "gpu_over_port": value / rpN,
                                             "e2e_over_reference_as_written": e2e_val / rateN, "e2e_over_port": e2e_val / rpN})
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
