"""Genealogy ingestion: host-side restatements of the R data-prep functions that feed the coalescent
likelihood (SURVEY.md section 8f.2 -- the "input format" row next to the hot path).

  coal_lik_init        R/coal_simulation.R:204-277   -> the `Init` sufficient statistics
  summarize_phylo      phylodyn (external, not in the reference tree): sampling / coalescent times of a tree
  volz_thin_sir        R/coal_simulation.R:11-70     -> synthetic genealogies by thinning (seeded NumPy RNG)

These run once per data set on the host (NumPy); the per-cell reduction of C*D and y happens on the
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


def volz_thin_sir(eff_pop, t_correct, samp_times, n_sampled, betaN, rng, lower=1.0):
    """R/coal_simulation.R:11-70: simulate coalescent times by thinning against the Volz rate on an SIR
    trajectory `eff_pop` (columns: time, S, I).  `rng` is a numpy Generator (replaces R's rexp/runif)."""
    eff_pop = np.asarray(eff_pop, dtype=np.float64)
    samp_times = np.asarray(samp_times, dtype=np.float64)
    n_sampled = np.asarray(n_sampled, dtype=np.int64)
    coal_times, lineages = [], []
    curr = 0
    active = int(n_sampled[curr])
    time = float(samp_times[curr])
    traj = 1.0 / (betaN * eff_pop[:, 1] / (eff_pop[:, 2] / 2.0))
    ts = t_correct - eff_pop[:, 0]
    i = int(np.argmin(ts >= 0)) - 1                # which.min(ts >= 0) - 1, then 0-based = that - 1 + ... see below
    # R: i = which.min(ts>=0) - 1 is a 1-based index of the last grid time <= t_correct; 0-based: subtract 1
    i = max(i, 0)
    max_samp = samp_times.max()
    while time <= max_samp or active > 1:
        if active == 1:
            curr += 1
            active += int(n_sampled[curr])
            time = float(samp_times[curr])
        time = time + rng.exponential(1.0 / (0.5 * active * (active - 1) / lower))
        while time > ts[i]:
            i -= 1
            if i < 0:
                i = 0
                break
        thed = traj[i]
        if curr < len(samp_times) - 1 and time >= samp_times[curr + 1]:
            curr += 1
            active += int(n_sampled[curr])
            time = float(samp_times[curr])
        elif rng.random() <= lower / thed:
            time = min(time, t_correct)
            coal_times.append(time)
            lineages.append(active)
            active -= 1
    coal_times = np.array(coal_times)
    return {"coal_times": coal_times, "lineages": np.array(lineages), "samp_times": samp_times,
            "n_sampled": n_sampled.astype(np.float64)}
