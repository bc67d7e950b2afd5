"""Do the speculative Metropolis pair and the two sequential Metropolis entries give the same chains at every batch size?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
def run(B, G, pair, split, iters=3):
    os.environ["LNA_CHAIN_GROUPS"] = str(G); os.environ["LNA_MH_PAIR"] = str(pair)
    if split is None: os.environ.pop("LNA_SPLIT", None)
    else: os.environ["LNA_SPLIT"] = str(split)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    out = []
    for it in range(iters):
        ch.step_rng(chm.vignette_opts(g, update_traj=it >= 1), 5, it)
        out.append(ch.get())
    return out
for B, G in ((12, 1), (2048, 1), (4096, 1), (4096, 2)):
    for split in (0, 1, None):
        a, b = run(B, G, 1, split), run(B, G, 0, split)
        msg = "same"
        for it, (x, y) in enumerate(zip(a, b)):
            for k in ("par", "coalLog", "n_full_evals", "n_mh_accept", "OriginTraj"):
                if not np.array_equal(x[k], y[k]):
                    bad = np.flatnonzero(np.any(np.atleast_2d(x[k].reshape(B, -1) != y[k].reshape(B, -1)), axis=1))
                    msg = f"DIFF at iteration {it} field {k}: {len(bad)} chains, first {bad[:5]}, evals pair {x['n_full_evals'][bad[0]]} seq {y['n_full_evals'][bad[0]]} acc {x['n_mh_accept'][bad[0]]}/{y['n_mh_accept'][bad[0]]}"
                    break
            if msg != "same": break
        print(f"B={B} G={G} LNA_SPLIT={split}: {msg}", flush=True)
