"""Experiment: three-warp pipelined candidate kernels forced on/off for the whole iteration, by chain count and groups."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
iters = 30
for B in (1024, 2048, 4096, 8192):
    for G in (1, 2):
        for split in (None, 0, 1):
            for plan in ("", "8:0:1,8:8:1,8:16:0"):
                os.environ["LNA_CHAIN_GROUPS"] = str(G)
                if split is None: os.environ.pop("LNA_SPLIT", None)
                else: os.environ["LNA_SPLIT"] = str(split)
                if plan: os.environ["LNA_SLICE_PLAN"] = plan
                else: os.environ.pop("LNA_SLICE_PLAN", None)
                ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
                ob, of = chm.vignette_opts(g, update_traj=False), chm.vignette_opts(g, update_traj=True)
                for it in range(4): ch.step_rng(ob, 1, it)
                for it in range(4, 10): ch.step_rng(of, 1, it)
                ch.sync(); t0 = time.perf_counter()
                for it in range(10, 10 + iters): ch.step_rng(of, 1, it)
                ch.sync(); dt = time.perf_counter() - t0
                print(f"B={B} G={G} LNA_SPLIT={split} plan={plan or 'auto':20s} {1e3*dt/iters:.3f} ms/iter {B*iters/dt/1e6:.2f} Mci/s", flush=True)
                del ch
