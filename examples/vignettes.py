#!/usr/bin/env python
"""The reference's three vignettes, run with many chains on one GPU.

    python examples/vignettes.py changepoint --chains 4096 --niter 2000     # vignettes/Changpoint_sim.Rmd:89-97
    python examples/vignettes.py constant    --chains 1024 --niter 2000     # vignettes/constant_sim.Rmd:87-95
    python examples/vignettes.py ebola       --chains 4096 --niter 500 [--rdata .../data/Ebola_Sierra_Leone2014.RData]

Each call is the vignette's own driver call (same arguments, same option lists -- including the Ebola vignette's unusual
Update_Param_each_Norm entries) with `n_chains` added; it prints the wall time, chain-iterations/s and posterior summaries
(R0(t) on the change-point grid, gamma, the hyper-parameter) pooled over chains after the first half of the run.
The coalescent data of the two simulation vignettes are the ones cached by the reference's knitr run
(tests/golden/*.npz); the Ebola tree is read from the reference's .RData file when given, else from tests/golden/ebola_tree.npz.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def cached(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = lambda k: z["setting." + k]
    coal = {k: s("Init.args." + k) for k in ("samp_times", "n_sampled", "coal_times")}
    prior = {k: s("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")}
    proposal = {k: s("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")}
    control = {"alpha": s("control.alpha"), "R0": float(s("control.R0")[0]), "gamma": float(s("control.gamma")[0]),
               "ch": s("control.ch"), "traj": s("control.traj"), "hyper": float(s("control.hyper")[0])}
    return coal, s("times"), float(s("t_correct")[0]), s("x_r"), prior, proposal, control


def run(which, chains, niter, seed, rdata_path=None):
    from lnaphylodyn_b200 import genealogy, mcmc, rdata
    if which == "changepoint":
        coal, times, tc, x_r, prior, proposal, control = cached("changepoint_sim")
        nch = len(x_r) - 1
        call = lambda: mcmc.General_MCMC_with_ESlice(
            coal, times, tc, x_r[0], gridsize=50, niter=niter, burn=0, thin=max(niter // 20, 1), changetime=x_r[1:],
            DEMS=("S", "I"), prior=prior, proposal=proposal, control=control, ESS_vec=(0, 1, 0, 0, 1, 1, 0), likelihood="volz",
            model="SIR", Index=(0, 1), nparam=2, method="admcmc",
            options={"joint": True, "PCOV": None, "beta": 0.95, "burn1": niter // 2, "parIdlist": {"c": 4, "d": nch + 5},
                     "isLoglist": {"c": 1, "d": 0}, "priorIdlist": {"c": 3, "d": 5}, "up": 1000000, "tune": 0.01, "hyper": False},
            n_chains=chains, seed=seed)
    elif which == "constant":
        coal, times, tc, x_r, prior, proposal, control = cached("constant_sim")
        nch = len(x_r) - 1
        call = lambda: mcmc.General_MCMC2(
            coal, times, tc, x_r[0], gridsize=50, niter=niter, burn=0, thin=max(niter // 20, 1), changetime=x_r[1:],
            DEMS=("S", "I"), prior=prior, proposal=proposal, control=control, updateVec=(1, 1, 1, 0, 0.5, 1, 1),
            likelihood="volz", model="SIR", Index=(0, 1), nparam=2, method="admcmc",
            options={"joint": True, "PCOV": None, "beta": 0.95, "burn1": 10, "parIdlist": {"a": 2, "b": 3, "c": 4, "d": 5 + nch},
                     "isLoglist": {"a": 1, "b": 0, "c": 1, "d": 0}, "priorIdlist": {"a": 1, "b": 2, "c": 3, "d": 5}, "up": 1000000},
            n_chains=chains, seed=seed)
    else:
        if rdata_path:
            t = rdata.read_phylo(rdata_path)
        else:
            z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
            t = {"edge": z["edge"], "edge_length": z["edge_length"], "ntip": int(z["ntip"])}
        coal = genealogy.summarize_phylo(t["edge"], t["edge_length"], t["ntip"])
        times = np.linspace(0, 1.36, 2001)
        cht = times[::50][1:-1]
        nch = len(cht)
        x_r = np.concatenate([[7e6], cht])
        call = lambda: mcmc.General_MCMC_with_ESlice(
            coal, times, 1.36, 7e6, gridsize=50, niter=niter, burn=0, thin=max(niter // 20, 1), changetime=cht, DEMS=("S", "I"),
            prior={"pop_pr": [1, 0.5, 1, 1], "R0_pr": [0.7, 0.5], "gamma_pr": [3, 0.1], "mu_pr": [3, 0.15], "hyper_pr": [3, 0.2]},
            proposal={"pop_prop": 0.2, "R0_prop": 0.05, "gamma_prop": 0.09, "mu_prop": 0.1, "hyper_prop": 0.2},
            control={"alpha": [1 / 7e6], "R0": 2.0, "gamma": 30.0, "ch": np.full(nch, 0.975), "traj": np.zeros((40, 2)), "hyper": 20.0},
            ESS_vec=(0, 1, 0, 0, 1, 1, 0), likelihood="volz", model="SIR", Index=(0, 1), nparam=2, method="admcmc",
            options={"joint": True, "PCOV": None, "beta": 0.95, "burn1": int(0.6 * niter), "parIdlist": {"a": 1, "c": 3, "d": nch + 5},
                     "isLoglist": {"a": 1, "c": 1, "d": 0}, "priorIdlist": {"a": 1, "c": 3, "d": 5}, "up": 1000000, "tune": 0.01,
                     "hyper": False},
            n_chains=chains, seed=seed)
    t0 = time.perf_counter()
    res = call()
    dt = time.perf_counter() - t0
    par = res["par"].reshape(chains, niter, -1)[:, niter // 2:, :]
    r0t = par[:, :, 2:3] * np.cumprod(par[:, :, 4:4 + nch], axis=2)  # R0 after each change point
    out = {"seconds": dt, "chain_iterations_per_s": chains * niter / dt,
           "R0_initial": (float(par[:, :, 2].mean()), float(par[:, :, 2].std())),
           "gamma": (float(par[:, :, 3].mean()), float(par[:, :, 3].std())),
           "hyper": (float(par[:, :, -1].mean()), float(par[:, :, -1].std())),
           "R0_t_mean": r0t.mean(axis=(0, 1)), "coalLog_mean": float(np.asarray(res["l1"]).reshape(chains, niter)[:, niter // 2:].mean())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["changepoint", "constant", "ebola"])
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--niter", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rdata", default=None)
    a = ap.parse_args()
    o = run(a.which, a.chains, a.niter, a.seed, a.rdata)
    np.set_printoptions(precision=3, linewidth=150, suppress=True)
    print(f"{a.which}: {a.chains} chains x {a.niter} iterations in {o['seconds']:.2f} s "
          f"({o['chain_iterations_per_s'] / 1e6:.2f} M chain-iterations/s incl. set-up, per-iteration records and thinned trajectories)")
    print(f"  initial R0 {o['R0_initial'][0]:.3f} +- {o['R0_initial'][1]:.3f}   gamma {o['gamma'][0]:.4f} +- {o['gamma'][1]:.4f}   "
          f"hyper {o['hyper'][0]:.2f} +- {o['hyper'][1]:.2f}   mean coalLog {o['coalLog_mean']:.2f}")
    try:
        print("  posterior mean R0(t) after each change point:", o["R0_t_mean"])
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
