"""Experiment (GPU box): General_MCMC2 iteration (constant_sim vignette) with one against two stream groups at 8192 / 16 384 chains
(measured: one group wins, 1.90 / 2.10 and 3.01 / 3.18 ms per iteration)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import _lib, api, chains as chm
g = Golden("constant_sim")
par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                       g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
opts = chm.make_mcmc2_opts(prior, proposal)
L = _lib.lib()
for B in (8192, 16384):
    for G in ("1", "2"):
        os.environ["LNA_CHAIN_GROUPS"] = G
        ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
        for it in range(6):
            _lib.check(L.lna_chains_step_mcmc2_rng(ch.h, C.byref(opts), 5, it, 0))
        ch.sync(); pb.sync(); t0 = time.perf_counter()
        n = 30
        for it in range(6, 6 + n):
            _lib.check(L.lna_chains_step_mcmc2_rng(ch.h, C.byref(opts), 5, it, 0))
        ch.sync(); pb.sync(); dt = (time.perf_counter() - t0) / n
        print(f"B={B:5d} G={G}: {1e3 * dt:.3f} ms per General_MCMC2 iteration, {B / dt / 1e6:.3f} M chain-iterations/s", flush=True)
