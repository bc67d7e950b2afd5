"""Histogram of the number of slice candidates ESlice_par_General needs (for choosing the phase widths)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm
g = Golden("changepoint_sim")
B = 4096
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
rng = np.random.default_rng(0)
prev = np.zeros(B, dtype=int)
hist = np.zeros(23, dtype=int)
for it in range(30):
    nrm = rng.standard_normal((B, ch.n_norm)); unf = rng.random((B, chm.ITER_UNIF))
    ch.step(chm.vignette_opts(g, update_traj=it >= 5), nrm, unf)
    s = ch.get()
    n = s["n_full_evals"] - prev - 2
    prev = s["n_full_evals"].copy()
    if it >= 10:
        hist += np.bincount(n, minlength=23)[:23]
p = hist / hist.sum()
print("candidates needed : share, cumulative")
for i in range(1, 22):
    print(i, f"{p[i]:.3f} {p[:i+1].sum():.3f}")
