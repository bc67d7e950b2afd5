import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api
g = Golden("changepoint_sim")
par = g.final["par"].ravel().copy(); eps = g.final["OriginTraj"].copy()
pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
for B in (1, 32, 1024, 4096, 14000):
    P = np.tile(par, (B, 1)); E = np.tile(eps, (B, 1, 1))
    out = {}
    for mode in ("0", "1"):
        os.environ["LNA_SPLIT"] = mode
        for _ in range(5): r = pb.full_loglik(P[:, 2:], P[:, :2], E, want=("coalLog",))
        t0 = time.perf_counter()
        for _ in range(30): r = pb.full_loglik(P[:, 2:], P[:, :2], E, want=("coalLog",))
        out[mode] = ((time.perf_counter() - t0) / 30, r["coalLog"].copy())
    same = np.array_equal(out["0"][1], out["1"][1])
    print(f"B={B:6d} host-call latency one-warp {1e6*out['0'][0]:8.1f} us  pipeline {1e6*out['1'][0]:8.1f} us  identical={same}")
