"""Stress of the three-warp pipeline's hand-over protocol: many FULL batches of varying size and many chain iterations,
pipeline vs one-warp kernels, must agree bit for bit every time (a lost update or a stale read would show as a mismatch, a
missed wake-up as a hang -- run under `timeout`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm
g = Golden("changepoint_sim")
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
rng = np.random.default_rng(2026)
par0 = g.final["par"].ravel()
t0 = time.time(); n_full = n_iter = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    B = int(rng.choice([1, 2, 31, 32, 33, 95, 160, 777, 2048, 4000]))
    par = np.tile(par0, (B, 1)); par[:, 2] *= np.exp(0.05 * rng.standard_normal(B)); par[:, 3] *= np.exp(0.05 * rng.standard_normal(B))
    eps = 0.5 * rng.standard_normal((B, 40, 2))
    out = {}
    for mode in ("0", "1"):
        os.environ["LNA_SPLIT"] = mode
        out[mode] = pb.full_loglik(par[:, 2:], par[:, :2], eps, want=("coalLog", "LatentTraj", "FT"))
    assert np.array_equal(out["0"]["coalLog"], out["1"]["coalLog"]) and np.array_equal(out["0"]["LatentTraj"], out["1"]["LatentTraj"])
    assert np.array_equal(out["0"]["FT"]["Sigma"], out["1"]["FT"]["Sigma"]), ("FULL", rep, B)
    n_full += B
    Bc = int(rng.choice([1, 7, 64, 300, 1000]))
    res = {}
    for mode in ("0", "1"):
        os.environ["LNA_SPLIT"] = mode
        ch = chm.Chains(pb, np.tile(np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                                                     g.setting("control.ch"), g.setting("control.hyper")]), (Bc, 1)), g.setting("control.traj"))
        for it in range(8):
            ch.step_rng(chm.vignette_opts(g, update_traj=it >= 2), 1000 + rep, it)
        res[mode] = ch.get()
    for k in ("par", "coalLog", "OriginTraj", "n_full_evals"):
        assert np.array_equal(res["0"][k], res["1"][k]), ("chains", rep, Bc, k)
    n_iter += 8 * Bc
print(f"ok: {n_full} FULL evaluations and {n_iter} chain-iterations compared bit for bit in {time.time() - t0:.1f} s")
