"""Small workload for compute-sanitizer: the three-warp pipelined kernels (k_full_loglik_split, k_cand_eval<*,*,1>) and the
one-warp kernels on a handful of chains, checked against each other.  usage: sanitize_probe.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm
os.environ["LNA_GRAPH"] = "0"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
res = {}
for split in ("1", "0"):
    os.environ["LNA_SPLIT"] = split
    rng = np.random.default_rng(0)
    B = 40
    par = np.tile(g.final["par"].ravel(), (B, 1)); par[:, 2] *= np.exp(0.05 * rng.standard_normal(B))
    eps = rng.standard_normal((B, 40, 2)) * 0.3
    r = pb.full_loglik(par[:, 2:], par[:, :2], eps, want=("coalLog", "LatentTraj", "FT"))
    ch = chm.Chains(pb, np.tile(par0, (12, 1)), g.setting("control.traj"))
    for it in range(iters):
        ch.step_rng(chm.vignette_opts(g, update_traj=it >= 1), 3, it)
    s = ch.get()
    res[split] = (r["coalLog"], r["LatentTraj"], s["par"], s["coalLog"])
for a, b in zip(res["1"], res["0"]):
    assert np.array_equal(a, b)
print("pipelined and one-warp kernels agree bit for bit; launches", pb.launch_count())
