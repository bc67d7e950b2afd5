"""Experiment: slice phase plans (LNA_SLICE_PLAN) x stream groups (LNA_CHAIN_GROUPS) for the resident-chain iteration."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time, numpy as np
    from conftest import Golden
    from lnaphylodyn_b200 import api, chains as chm
    B, iters = int(sys.argv[2]), 12
    g = Golden("changepoint_sim")
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    free = os.environ.get("LNA_FREE", "0") != "0"   # stream groups not joined after every step, started staggered
    iters = int(os.environ.get("LNA_ITERS", iters))
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=free)
    opts = chm.vignette_opts(g, update_traj=True)
    for it in range(6): ch.step_rng(opts, 1, it)
    ch.sync(); pb.sync(); t0 = time.perf_counter()
    for it in range(6, 6 + iters): ch.step_rng(opts, 1, it)
    ch.sync(); pb.sync(); dt = time.perf_counter() - t0
    s = ch.get()
    print(f"{1e3*dt/iters:.3f} ms/iter {B*iters/dt/1e6:.2f} M chain-iter/s  evals/iter {s['n_full_evals'].mean()/(6+iters):.2f} coalLog {s['coalLog'].mean():.3f}")
    sys.exit(0)
B = sys.argv[1] if len(sys.argv) > 1 else "4096"
plans = ["", "8:0:1,16:8:0", "8:0:1,8:8:1,8:16:0", "8:0:1,4:8:1,4:12:1,8:16:0", "8:0:1,4:8:1,8:12:0", "8:0:1,4:8:1,16:12:0", "4:0:1,4:4:1,4:8:1,4:12:1,8:16:0",
         "4:0:2,4:8:1,4:12:1,8:16:0", "8:0:1,2:8:1,2:10:1,4:12:1,8:16:0"]
for G in ("1", "2", "4"):
    for plan in plans:
        e = dict(os.environ, LNA_CHAIN_GROUPS=G)
        if plan: e["LNA_SLICE_PLAN"] = plan
        else: e.pop("LNA_SLICE_PLAN", None)
        r = subprocess.run([sys.executable, __file__, "child", B], env=e, capture_output=True, text=True)
        print(f"G={G} plan={plan or 'auto':40s} {r.stdout.strip().splitlines()[-1] if r.returncode == 0 else 'FAILED ' + r.stderr[-300:]}")
