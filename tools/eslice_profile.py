"""Run a few resident-chain iterations (for ncu launch lists / profiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
rng = np.random.default_rng(0)
for it in range(iters):
    nrm = rng.standard_normal((B, ch.n_norm)); unf = rng.random((B, chm.ITER_UNIF))
    ch.step(chm.vignette_opts(g, update_traj=it >= 1), nrm, unf)
s = ch.get()
print("mean coalLog", s["coalLog"].mean(), "full evals/iter", s["n_full_evals"].mean() / iters, "cheap", s["n_cheap_evals"].mean() / iters)
