"""Genealogy ingestion: host-side restatements of the R data-prep functions that feed the coalescent
likelihood (SURVEY.md section 8f.2 -- the "input format" row next to the hot path).

  coal_lik_init        R/coal_simulation.R:204-277   -> the `Init` sufficient statistics
  summarize_phylo      phylodyn (external, not in the reference tree): sampling / coalescent times of a tree
  volz_thin_sir        R/coal_simulation.R:11-70     -> synthetic genealogies by thinning, ON THE DEVICE (csrc/lna_thin.cuh)

coal_lik_init / summarize_phylo run once per data set on the host (NumPy); the per-cell reduction of C*D and y happens on the
device when the problem is created (csrc/lna_kernels.cuh: k_cell_sums).
"""
import numpy as np


def coal_lik_init(samp_times, n_sampled, coal_times, grid):
    """R/coal_simulation.R:204-277.  Returns the entries of `Init` the likelihoods read (by name)."""
    samp_times = np.asarray(samp_times, dtype=np.float64)
    n_sampled = np.asarray(n_sampled, dtype=np.float64)
    coal_times = np.asarray(coal_times, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    ns, nc = len(samp_times), len(coal_times)
    Tcoal = coal_times.max()
    if Tcoal > grid.max():
        raise ValueError("coalescent time out of grid")
    n0 = int(np.argmin(grid < Tcoal)) + 1          # which.min(grid < Tcoal), 1-based
    grid = grid[:n0]
    ng = n0
    if len(samp_times) != len(n_sampled):
        raise ValueError("samp_times vector of differing length than n_sampled vector.")
    t = np.unique(np.concatenate([samp_times, coal_times, grid]))
    l = np.zeros(len(t))
    for i in range(ns):
        l[t >= samp_times[i]] += n_sampled[i]
    # l[t >= coal_times[i]] -= 1 for every coalescence = minus the number of coalescent times <= t
    l -= np.searchsorted(np.sort(coal_times), t, side="right")
    if np.any((l < 1) & (t >= samp_times.min())):
        raise ValueError("Number of active lineages falls below 1 after the first sampling point.")
    mask = l > 0
    t = t[mask]
    l = l[mask][:-1]
    gridrep = np.zeros(ng)
    for i in range(ng - 1):
        gridrep[i] = np.sum((t > grid[i]) & (t <= grid[i + 1]))
    # gridrep[ng-1] compares against grid[ng+1] = NA in R: sum(FALSE & NA) = 0 because no t exceeds grid[ng]
    C = 0.5 * l * (l - 1)
    D = np.diff(t)
    y = np.isin(t[1:], coal_times).astype(np.float64)
    return {"t": t, "l": l, "C": C, "D": D, "y": y, "gridrep": gridrep, "ns": float(n_sampled.sum()), "nc": nc,
            "ng": ng, "args": {"samp_times": samp_times, "n_sampled": n_sampled, "coal_times": coal_times,
                               "grid": grid}}


def summarize_phylo(edge, edge_length, ntip):
    """What phylodyn::summarize_phylo returns for an ape `phylo` (restated; the package is not in the
    reference tree): node heights measured from the most recent tip; samp_times = sorted unique tip
    heights with multiplicities n_sampled; coal_times = sorted internal-node heights."""
    edge = np.asarray(edge, dtype=np.int64)
    edge_length = np.asarray(edge_length, dtype=np.float64)
    nnode_total = int(edge.max())
    parent = np.zeros(nnode_total + 1, dtype=np.int64)
    blen = np.zeros(nnode_total + 1)
    parent[edge[:, 1]] = edge[:, 0]
    blen[edge[:, 1]] = edge_length
    root = ntip + 1
    depth = np.full(nnode_total + 1, np.nan)
    depth[root] = 0.0
    # ape stores edges in an order where parents are not guaranteed first: resolve iteratively
    order = list(range(1, nnode_total + 1))
    pending = [v for v in order if v != root]
    while pending:
        nxt = []
        for v in pending:
            if np.isnan(depth[parent[v]]):
                nxt.append(v)
            else:
                depth[v] = depth[parent[v]] + blen[v]
        if len(nxt) == len(pending):
            raise ValueError("tree is not connected to the root")
        pending = nxt
    height = np.nanmax(depth[1:]) - depth
    tips = height[1:ntip + 1]
    internal = height[ntip + 1:nnode_total + 1]
    samp_times, n_sampled = np.unique(tips, return_counts=True)
    return {"samp_times": samp_times, "n_sampled": n_sampled.astype(np.float64), "coal_times": np.sort(internal)}


def volz_thin_sir(eff_pop, t_correct, samp_times, n_sampled, betaN, seed=0, lower=1.0, n_genealogies=None, item_offset=0,
                  max_draws=0, device=0):
    """R/coal_simulation.R:11-70 on the device (csrc/lna_thin.cuh: one warp per genealogy, 32 thinning proposals per round):
    coalescent times simulated by thinning against the Volz rate on SIR trajectories `eff_pop` (columns time, S, I; one
    n x 3 trajectory, or B x n x 3 for B genealogies on their own trajectories) with `betaN` on the same grid.  R's rexp /
    runif are a counter-based stream keyed by `seed` and indexed by (item_offset + genealogy, proposal number).
    `n_genealogies = B` with a single trajectory draws B independent genealogies on it; the result is then a list."""
    from . import _lib, api
    E = np.asarray(eff_pop, dtype=np.float64)
    single_traj = E.ndim == 2
    B = int(n_genealogies) if n_genealogies is not None else (1 if single_traj else E.shape[0])
    if not single_traj and E.shape[0] != B:
        raise ValueError("n_genealogies must match the number of trajectories")
    n = E.shape[-2]
    if not np.all(E[..., -1, 0] > t_correct):
        # R: i = which.min(ts >= 0) - 1 is 0 when no grid time lies beyond t_correct, and `ts[0]` then stops the loop with an error
        raise ValueError("volz_thin_sir: the time grid must extend beyond t_correct")
    cols = [np.ascontiguousarray(E[..., j]).reshape(-1) for j in range(3)]
    bn = np.ascontiguousarray(np.broadcast_to(np.asarray(betaN, dtype=np.float64), E.shape[:-1])).reshape(-1)
    st = np.ascontiguousarray(np.asarray(samp_times, dtype=np.float64))
    ns = np.ascontiguousarray(np.asarray(n_sampled, dtype=np.float64).round().astype(np.int32))
    max_coal = int(ns.sum()) - 1
    pb = api.problem_for(np.array([0.0, 1.0]), [1.0], [0, 2, 0, 1], 1, device=device)
    ct, lin = np.empty((B, max_coal)), np.empty((B, max_coal), dtype=np.int32)
    nc, stt, nd = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.uint64)
    _lib.check(_lib.lib().lna_volz_thin_sir(pb.h, B, n, 0 if single_traj else n, api._ptr(cols[0]), api._ptr(cols[1]),
                                            api._ptr(cols[2]), api._ptr(bn), len(st), api._ptr(st), api._ptr(ns),
                                            float(t_correct), float(lower), int(seed), int(item_offset), int(max_draws),
                                            max_coal, api._ptr(ct), api._ptr(lin), api._ptr(nc), api._ptr(nd), api._ptr(stt)),
               "lna_volz_thin_sir")
    out = [{"coal_times": ct[b, :nc[b]].copy(), "lineages": lin[b, :nc[b]].copy(), "samp_times": st,
            "n_sampled": ns.astype(np.float64), "n_draws": int(nd[b]), "status": int(stt[b])} for b in range(B)]
    for o in out:
        o["intercoal_times"] = np.diff(o["coal_times"], prepend=0.0)
    return out[0] if (n_genealogies is None and single_traj) else out


def thin_draws(seed, item, first, count, device=0):
    """proposals first .. first + count - 1 of genealogy `item`'s stream: (waiting-time uniforms, thinning marks)"""
    from . import _lib, api
    pb = api.problem_for(np.array([0.0, 1.0]), [1.0], [0, 2, 0, 1], 1, device=device)
    ua, ub = np.empty(count), np.empty(count)
    _lib.check(_lib.lib().lna_thin_draws(pb.h, int(seed), int(item), int(first), int(count), api._ptr(ua), api._ptr(ub)),
               "lna_thin_draws")
    return ua, ub
