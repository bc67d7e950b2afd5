"""Experiment: free-running, staggered stream groups x phase plans x Metropolis policy for the resident-chain iteration
(step_rng).  usage: eslice_free.py [B] [iters]; LNA_GRAPH=0 for eager launches."""
import os, sys, time, hashlib, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
plans = ["", "8:0:1,8:8:1,8:16:0", "8:0:1,4:8:1,4:12:1,8:16:0"]
ref = None
print(f"B={B} iters={iters} LNA_GRAPH={os.environ.get('LNA_GRAPH','1')}")
def run(free, G, plan, stag, pair):
    global ref
    os.environ["LNA_CHAIN_GROUPS"] = str(G)
    os.environ["LNA_STAGGER_US"] = str(stag)
    if pair is None: os.environ.pop("LNA_MH_PAIR", None)
    else: os.environ["LNA_MH_PAIR"] = str(pair)
    if plan: os.environ["LNA_SLICE_PLAN"] = plan
    else: os.environ.pop("LNA_SLICE_PLAN", None)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=bool(free))
    ob, of = chm.vignette_opts(g, update_traj=False), chm.vignette_opts(g, update_traj=True)
    for it in range(4): ch.step_rng(ob, 1, it)
    for it in range(4, 12): ch.step_rng(of, 1, it)
    ch.sync(); s0 = ch.get()
    if free: ch.set_free_running(False); ch.set_free_running(True)   # re-stagger after the synchronising get()
    t0 = time.perf_counter()
    for it in range(12, 12 + iters): ch.step_rng(of, 1, it)
    ch.sync(); dt = time.perf_counter() - t0
    s = ch.get()
    h = hashlib.sha1(s["par"].tobytes() + s["coalLog"].tobytes() + s["OriginTraj"].tobytes()).hexdigest()[:10]
    if ref is None: ref = h
    ne = (s["n_full_evals"] - s0["n_full_evals"]).mean() / iters
    print(f"free={free} G={G} stagger={stag:6.1f} pair={pair} plan={plan or 'auto':28s} {1e3*dt/iters:.3f} ms/iter {B*iters/dt/1e6:.2f} Mci/s evals/iter {ne:.2f} {'OK' if h == ref else 'MISMATCH ' + h}", flush=True)
    del ch
run(0, 1, "", 0, None)
run(0, 2, "", 0, None)
for G in (2, 3, 4, 6, 8):
    for stag in (0.0, 350.0 / G, 700.0 / G, 1000.0 / G):
        for plan in plans:
            for pair in (None, 0):
                run(1, G, plan, stag, pair)
