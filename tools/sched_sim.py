"""Offline model of the barrier-free slice kernel's schedule (CPU oracle): for a few hundred slice updates of the golden chain
state, every candidate's outcome (accepted, or the LNA interval at which its rejection is proven), then the number of ticks a
chain / a warp of 4 chains needs under different lane-assignment rules.  usage: python tools/sched_sim.py [updates]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from conftest import Golden
from oracle import oracle as orc
g = Golden("changepoint_sim")
par = g.final["par"].ravel().copy(); eps = g.final["OriginTraj"].copy()
ini = g.init; C, D, y, gridrep = ini["C"], ini["D"], ini["y"], ini["gridrep"].astype(int); ng = int(np.ravel(ini["ng"])[0]); n0 = ng - 1
start = np.concatenate([[0], np.cumsum(gridrep)])
cellCD = np.array([np.sum(C[start[i]:start[i+1]] * D[start[i]:start[i+1]]) for i in range(n0)])
cellY = np.array([np.sum(y[start[i]:start[i+1]]) for i in range(n0)])
ub = np.where(cellY > 0, cellY * np.log(np.maximum(cellY, 1e-300) / (2 * cellCD)) - cellY, 0.0)
PRIORS = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)]; ESS = [0, 1, 0, 0, 1, 1, 0]
K = 40
def cells(parv):
    r = orc.New_Param_List(parv[2:], parv[:2], g.gridsize, g.times, g.x_r, g.x_i)
    lat = orc.TransformTraj(r["Ode"], eps, r["FT"]); bet = r["betaN"]
    L = np.argmax(lat[:, 0] >= g.t_correct)
    t = np.zeros(n0)
    for i in range(n0):
        row = L - i; S, I = lat[row, 1], lat[row, 2]
        if S <= 0 or I <= 0: return None, L
        lam = bet[row] * S / I
        t[i] = -2 * cellCD[i] * lam + cellY[i] * np.log(lam)
    return t, L
cur, L = cells(par); cl = cur.sum()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(0)
updates = []   # per update: list over candidates 1..21 of ticks (negative = accepted after 40 ticks)
for rep in range(N):
    nrm = rng.standard_normal(43); u = rng.uniform(size=23)
    logy = cl + np.log(u[0])
    th = 2 * np.pi * u[1]; lo, hi = th - 2 * np.pi, th
    out = []; first = None
    for j in range(1, 21):
        newpar = orc.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nrm, ESS, PRIORS)
        try: t, _ = cells(newpar)
        except Exception: t = None
        if t is None: out.append(1)
        else:
            ll = t.sum()
            if ll > logy:
                out.append(-K)
                if first is None: first = j
            else:
                order = np.arange(n0 - 1, -1, -1)
                part = np.cumsum(t[order]); rem = np.concatenate([np.cumsum(ub[order][::-1])[::-1][1:], [0.0]])
                ok = np.nonzero(part + rem < logy - 1e-6)[0]
                row = (L - order[ok[0]]) if len(ok) else K
                out.append(max(1, min(K, int(row))))
        if first is not None and j >= first + 10: break
        if th < 0: lo = th
        else: hi = th
        th = lo + (hi - lo) * u[j + 1]
    while len(out) < 21: out.append(-K)   # the cap candidate (always accepted) / unexplored tail
    updates.append((first if first else 21, out))
acc = np.array([a for a, _ in updates]); print("updates", N, "mean accepted", acc.mean(), "hist", np.bincount(acc, minlength=22)[1:])

def simulate(out, accepted, W, order_fn):
    """ticks until the chain is resolved.  order_fn(W) -> the order in which candidates are started."""
    order = order_fn(W)
    lane_free = [0] * W; done = {}; started = 0; t_acc = None; best = 10**9
    # event simulation by ticks
    running = {}  # lane -> (cand, end_tick)
    t = 0
    while True:
        # finish
        for ln in list(running):
            c, e = running[ln]
            if e <= t:
                done[c] = e; del running[ln]
                if out[c - 1] < 0: best = min(best, c); lane_free[ln] = 10**9   # accepted: the lane stops
        if best < 10**9 and all((c in done) for c in range(1, best)) :
            return t
        for ln in range(W):
            if ln not in running and lane_free[ln] <= t:
                while started < len(order) and (order[started] in done or order[started] >= best): started += 1
                if started < len(order):
                    c = order[started]; started += 1
                    running[ln] = (c, t + abs(out[c - 1]))
        t += 1
        if t > 400: return t

def natural(W): return list(range(1, 22))
def shifted(s):
    def f(W):
        first = list(range(s, s + W)); rest = [c for c in range(1, 22) if c not in first]
        return first + rest
    return f
def static_ticks(out, accepted, W):
    # lane w: candidates w+1, w+1+W, ...
    end = {}
    for w in range(W):
        t = 0
        for c in range(w + 1, 22, W):
            if c > accepted: break
            t += abs(out[c - 1]); end[c] = t
            if out[c - 1] < 0: break
    return max(end[c] for c in range(1, accepted + 1))
for W in (8, 16):
    res = {"static": [static_ticks(o, a, W) for a, o in updates], "dynamic": [simulate(o, a, W, natural) for a, o in updates]}
    for s in (2, 3, 4, 5, 6):
        res[f"dynamic, first wave {s}..{s+W-1}"] = [simulate(o, a, W, shifted(s)) for a, o in updates]
    for name, v in res.items():
        v = np.array(v)
        # warp = 4 chains: max
        r4 = v[: len(v) // 4 * 4].reshape(-1, 4).max(axis=1)
        print(f"W={W:2d} {name:32s} chain ticks mean {v.mean():5.1f} p90 {np.percentile(v, 90):5.1f} max {v.max():3d} | warp(4 chains) mean {r4.mean():5.1f} max {r4.max():3d}")
