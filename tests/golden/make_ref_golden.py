"""Generate tests/golden/ref_as_written.npz: inputs and outputs of THE REFERENCE'S OWN CODE (compiled verbatim from
/root/reference/src on the stand-in headers of oracle/shim/, see oracle/ref_harness.cpp) for the hot-path functions,
on fixed seeded inputs.  The GPU box has no /root/reference; tests/test_gpu_vs_reference.py compares the CUDA path
with these vectors there.

    python tests/golden/make_ref_golden.py        (needs /root/reference; ~20 s)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden  # noqa: E402
from oracle import reference as ref  # noqa: E402

PRIORS = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)]
ESS = [0, 1, 0, 0, 1, 1, 0]
out = {}


def perturbed(g, B, seed, scale):
    rng = np.random.default_rng(seed)
    par = np.tile(g.final["par"].ravel(), (B, 1))
    par[1:, 2] *= np.exp(scale * rng.standard_normal(B - 1))
    par[1:, 3] *= np.exp(scale * rng.standard_normal(B - 1))
    par[1:, 4:-1] *= np.exp(0.2 * scale * rng.standard_normal((B - 1, par.shape[1] - 5)))
    eps = np.tile(g.final["OriginTraj"], (B, 1, 1))
    eps[1:] += scale * rng.standard_normal((B - 1,) + eps.shape[1:])
    return par, eps


def full_evals(tag, g, B, seed, scale):
    par, eps = perturbed(g, B, seed, scale)
    cfg, ini = ref.Cfg(g.x_r, g.x_i, par.shape[1] - 2), ref.Init(g.init)
    res = [ref.full_loglik(cfg, ini, par[b, 2:], par[b, :2], g.times, g.gridsize, eps[b], g.t_correct) for b in range(B)]
    out[tag + ".par"], out[tag + ".eps"] = par, eps
    out[tag + ".status"] = np.array([r["status"] for r in res], dtype=np.int32)
    for k in ("coalLog", "Ode", "betaN", "LatentTraj"):
        out[f"{tag}.{k}"] = np.array([r[k] for r in res])
    out[tag + ".A"] = np.array([r["FT"]["A"] for r in res])
    out[tag + ".L"] = np.array([r["FT"]["Sigma"] for r in res])


def main():
    cp, co = Golden("changepoint_sim"), Golden("constant_sim")
    full_evals("full_cp", cp, 48, 1, 0.05)
    full_evals("full_co", co, 24, 2, 0.05)

    # SEIR, p = 3
    g = cp
    rng = np.random.default_rng(21)
    B, x_i = 16, [39, 3, 0, 2]
    param = np.empty((B, 43))
    param[:, 0] = 2.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 1] = 0.5 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 2] = 0.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 3:42] = np.exp(0.02 * rng.standard_normal((B, 39)))
    param[:, 42] = 20.0
    initial = np.tile([1e6, 5.0, 5.0], (B, 1))
    eps = 0.5 * rng.standard_normal((B, 40, 3))
    cfg, ini = ref.Cfg(g.x_r, x_i, 43, model="SEIR"), ref.Init(g.init)
    res = [ref.full_loglik(cfg, ini, param[b], initial[b], g.times, g.gridsize, eps[b], g.t_correct) for b in range(B)]
    out["seir.param"], out["seir.initial"], out["seir.eps"] = param, initial, eps
    out["seir.coalLog"] = np.array([r["coalLog"] for r in res])
    out["seir.LatentTraj"] = np.array([r["LatentTraj"] for r in res])

    # Update_Param decisions
    par, _ = perturbed(g, 32, 5, 0.01)
    u = np.random.default_rng(9).uniform(size=32)
    cl0 = float(g.final["coalLog"][0])
    res = [ref.Update_Param(par[b, 2:], par[b, :2], g.times, g.final["OriginTraj"], g.x_r, g.x_i, g.init, g.gridsize, cl0, 0.1,
                            g.t_correct, unif=u[b]) for b in range(32)]
    out["mh.par"], out["mh.unif"] = par, u
    out["mh.accept"] = np.array([int(bool(r["accept"])) for r in res], dtype=np.int32)
    out["mh.coalLog"] = np.array([r["coalLog"] if r["accept"] else np.nan for r in res])

    # ESlice_par_General from the cached final state
    st_par, st_eps = g.final["par"].ravel(), g.final["OriginTraj"]
    B = 32
    rng = np.random.default_rng(100)
    nrm, unf = rng.standard_normal((B, 43)), rng.random((B, 22))
    unf[:4, 0] = 1.0 - 1e-9 * rng.random(4)
    res = [ref.ESlice_par_General(st_par, g.times, st_eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, cl0, g.t_correct,
                                  normals=nrm[b], uniforms=unf[b]) for b in range(B)]
    out["esp.normals"], out["esp.uniforms"] = nrm, unf
    out["esp.n_uniforms"] = np.array([r["n_uniforms"] for r in res], dtype=np.int32)
    for k in ("par", "CoalLog", "LatentTraj", "logOrigin", "betaN", "OdeTraj"):
        out["esp." + k] = np.array([r[k] for r in res])

    # ESlice_general_NC from the cached final state
    npl = ref.New_Param_List(st_par[2:], st_par[:2], g.gridsize, g.times, g.x_r, g.x_i)
    rng = np.random.default_rng(200)
    nrm, unf = rng.standard_normal((B, 80)), rng.random((B, 22))
    unf[:3, 0] = 1.0 - 1e-12
    res = [ref.ESlice_general_NC(st_eps, npl["Ode"], npl["FT"], st_par[:2], g.init, npl["betaN"], g.t_correct, 1.0, cl0,
                                 g.gridsize, True, "SIR", normals=nrm[b], uniforms=unf[b]) for b in range(B)]
    out["nc.normals"], out["nc.uniforms"] = nrm, unf
    out["nc.n_uniforms"] = np.array([r["n_uniforms"] for r in res], dtype=np.int32)
    for k in ("CoalLog", "LatentTraj", "logOrigin", "OriginTraj"):
        out["nc." + k] = np.array([r[k] for r in res])

    # ESlice_change_points (constant_sim)
    g2 = co
    p2, e2 = g2.final["par"].ravel(), g2.final["OriginTraj"]
    cl2 = float(g2.final["coalLog"][0])
    B2 = 16
    rng = np.random.default_rng(300)
    nrm, unf = rng.standard_normal((B2, 39)), rng.random((B2, 22))
    unf[:2, 0] = 1.0 - 1e-10
    res = [ref.ESlice_change_points(p2[2:], p2[:2], g2.times, e2, g2.x_r, g2.x_i, g2.init, g2.gridsize, cl2, g2.t_correct,
                                    normals=nrm[b], uniforms=unf[b]) for b in range(B2)]
    out["chp.normals"], out["chp.uniforms"] = nrm, unf
    out["chp.n_uniforms"] = np.array([r["n_uniforms"] for r in res], dtype=np.int32)
    for k in ("param", "CoalLog", "LatentTraj", "LogChProb"):
        out["chp." + k] = np.array([r[k] for r in res])

    # resident chains: 6 chains x 8 iterations of the General_MCMC_with_ESlice body, from the vignette's initial state
    rdriver = ref.driver()
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    Bc, iters = 6, 8
    rng = np.random.default_rng(400)
    nrm, unf = rng.standard_normal((iters, Bc, 43 + 80)), rng.random((iters, Bc, 47))
    states = [rdriver.init_state(g, par0, g.setting("control.traj")) for _ in range(Bc)]
    tr_par, tr_cl, tr_eps, tr_lat, tr_acc = (np.empty((iters, Bc, 44)), np.empty((iters, Bc)), np.empty((iters, Bc, 40, 2)),
                                             np.empty((iters, Bc, 41, 3)), np.empty((iters, Bc), dtype=np.int32))
    for it in range(iters):
        opts = rdriver.opts_from_golden(g, update_traj=it >= 2)
        for c in range(Bc):
            states[c] = rdriver.eslice_iteration(states[c], g, opts, nrm[it, c], unf[it, c])
            tr_par[it, c], tr_cl[it, c] = states[c]["par"], states[c]["coalLog"]
            tr_eps[it, c], tr_lat[it, c], tr_acc[it, c] = states[c]["OriginTraj"], states[c]["LatentTraj"], states[c]["n_mh_accept"]
    out["chain.par0"], out["chain.normals"], out["chain.uniforms"] = par0, nrm, unf
    out["chain.par"], out["chain.coalLog"], out["chain.OriginTraj"], out["chain.LatentTraj"] = tr_par, tr_cl, tr_eps, tr_lat
    out["chain.n_mh_accept"] = tr_acc

    path = os.path.join(HERE, "ref_as_written.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
