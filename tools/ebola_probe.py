"""configs[2]: the Ebola vignette's iteration AS WRITTEN (entries S0 / R0 / no-op hyper, parameter and trajectory slice updates)
on resident chains, 1024 chains per GPU (8192 over 8 GPUs): ms per iteration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lnaphylodyn_b200 import api, chains as chm, genealogy as gen
z = np.load(os.path.join(ROOT, "tests", "golden", "ebola_tree.npz"))
s = gen.summarize_phylo(z["edge"], z["edge_length"], int(z["ntip"]))
times = np.linspace(0, 1.36, 2001)
init = gen.coal_lik_init(s["samp_times"], s["n_sampled"], s["coal_times"], times[::50])
cht = times[::50][1:-1]; nch = len(cht)
pb = api.Problem(times, np.concatenate([[7e6], cht]), [nch, 2, 0, 1], 50, 1.36, init)
prior = [[1, 0.5, 1, 1], [0.7, 0.5], [3, 0.1], [3, 0.15], [3, 0.2]]
proposal = [[0.2], [0.05], [0.09], [0.1], [0.2]]
par0 = np.concatenate([[7e6, 1.0, 2.0, 30.0], np.full(nch, 0.975), [20.0]])
for tree in ("0", "1"):
    os.environ["LNA_MH_TREE"] = tree
    for B in (1024, 4096):
        ch = chm.Chains(pb, np.tile(par0, (B, 1)), np.zeros((40, 2)))
        opts = chm.make_opts_general(prior, proposal, [0, 1, 0, 0, 1, 1, 0], ((1, 1, 1), (3, 1, 3), (nch + 5, 0, 5)), hyper=False)
        for it in range(10): ch.step_rng(opts, 3, it)
        pb.sync(); t0 = time.perf_counter(); n = 60
        for it in range(10, 10 + n): ch.step_rng(opts, 3, it)
        pb.sync(); dt = (time.perf_counter() - t0) / n
        print(f"tree={tree} B={B}: {1e3 * dt:.3f} ms per iteration, {B / dt / 1e6:.2f} M chain-iterations/s, E = {len(init['C'])} events")
