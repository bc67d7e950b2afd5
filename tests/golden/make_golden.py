"""Generate the golden fixtures under tests/golden/ from the reference's own shipped artefacts.

Run ONCE in the build container (where /root/reference is mounted, read-only):

    python tests/golden/make_golden.py

Sources -- nothing is computed here, values are copied verbatim:
  * vignettes/Changpoint_sim_cache/html/MCMC_5186...rdb   -> changepoint_sim.npz  (`resres5`,
    output of General_MCMC_with_ESlice, R/LNA_chpt_slice.R:723-849)
  * vignettes/constant_sim_cache/markdown_github/MCMC_f493...rdb -> constant_sim.npz (`resres6`,
    output of General_MCMC2, R/LNA_chpt_slice.R:508-693)
  * data/Ebola_Sierra_Leone2014.RData -> ebola_tree.npz (ape::phylo edge list)

Each cache holds the MCMC settings (incl. the coal_lik_init `Init` list), the final chain state
`MCMC_obj` (par, OriginTraj, FT$A, FT$Sigma(=L), Ode_Traj_coarse, LatentTraj, betaN, coalLog, ...)
and 2000 per-iteration (par, Trajectory, coalLog) triplets.  To keep the fixtures small we keep
every 20th iteration (100 triplets per cache) plus the first 4.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from lnaphylodyn_b200 import rdata as rxdr  # noqa: E402  (the package's XDR reader)

REF = "/root/reference"


def _flat(prefix, d, out):
    for k, v in d.items():
        key = f"{prefix}{k}"
        if isinstance(v, dict):
            _flat(key + ".", v, out)
        elif isinstance(v, np.ndarray):
            out[key] = v
        elif isinstance(v, list) and v and isinstance(v[0], str):
            out[key] = np.array(v)
        elif v is None:
            continue
        else:
            raise TypeError((key, type(v)))


def cache_to_npz(rdb, var, dst, stride=20):
    res = rxdr.read_rdb(rdb, rdb[:-4] + ".rdx")[var]
    out = {}
    _flat("setting.", res["MCMC_setting"], out)
    _flat("final.", res["MCMC_obj"], out)
    niter = res["par"].shape[0]
    idx = np.unique(np.concatenate([np.arange(4), np.arange(stride - 1, niter, stride)]))
    out["iter.index"] = idx.astype(np.int64)          # 0-based MCMC iteration
    out["iter.par"] = res["par"][idx]                  # (m, 44)
    out["iter.Trajectory"] = np.ascontiguousarray(np.moveaxis(res["Trajectory"][:, :, idx], 2, 0))  # (m, 41, 3)
    out["iter.coalLog"] = res["l1"][idx]
    out["iter.logOrigin"] = res["l"][idx]
    out["iter.par_probs"] = res["l2"][idx]
    out["iter.chprobs"] = res["l3"][idx]
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst), "bytes;", len(out), "arrays;", len(idx), "iterations")


def main():
    cache_to_npz(f"{REF}/vignettes/Changpoint_sim_cache/html/MCMC_518601a265bb7a08541cb5357f32eb40.rdb",
                 "resres5", f"{HERE}/changepoint_sim.npz")
    cache_to_npz(f"{REF}/vignettes/constant_sim_cache/markdown_github/MCMC_f493722c202b7d35d3f27478a55da1b3.rdb",
                 "resres6", f"{HERE}/constant_sim.npz")
    tree = rxdr.read_rdata(f"{REF}/data/Ebola_Sierra_Leone2014.RData")["Ebola_Sierra_Leone2014"]
    dst = f"{HERE}/ebola_tree.npz"
    np.savez_compressed(dst, edge=tree["edge"].astype(np.int32), edge_length=tree["edge.length"],
                        Nnode=tree["Nnode"].astype(np.int32), ntip=np.int32(len(tree["tip.label"])))
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
