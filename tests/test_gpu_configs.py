"""Parity at the other BASELINE.json configs: the Ebola genealogy (configs[2]), a synthetic 5000-tip
genealogy on the 2000-step grid (configs[4]) and i.i.d. SEIR parameter draws (configs[3]); plus
size-independent properties at full batch sizes (chains are independent: permutation equivariance,
determinism, batch-size invariance)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from lnaphylodyn_b200 import genealogy as gen
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    import lnaphylodyn_b200 as m
    from lnaphylodyn_b200 import _lib
    assert _lib.lib().lna_device_count() > 0
    return m.api


def check_batch(api, times, x_r, x_i, init, t_correct, param, initial, eps, model="SIR", min_ok=4):
    pb = api.problem_for(times, x_r, x_i, 50, t_correct, init, model=model)
    res = pb.full_loglik(param, initial, eps, want=("coalLog", "LatentTraj", "betaN"))
    cfg, ini = orc.Cfg(x_r, x_i, param.shape[1], model=model), orc.Init(init)
    ok = 0
    for b in range(param.shape[0]):
        o = orc.full_loglik(cfg, ini, param[b], initial[b], times, 50, eps[b], t_correct)
        assert o["status"] == res["status"][b], (b, o["status"], res["status"][b])
        if o["coalLog"] == orc.SENTINEL:
            assert res["coalLog"][b] == orc.SENTINEL
            continue
        ok += 1
        assert res["coalLog"][b] == pytest.approx(o["coalLog"], rel=TOL), b
        assert relerr(res["betaN"][b], o["betaN"]) < TOL
        assert np.max(np.abs(res["LatentTraj"][b][:, 1:] - o["LatentTraj"][:, 1:]) /
                      np.maximum(np.abs(o["LatentTraj"][:, 1:]), 1.0)) < TOL
    assert ok >= min_ok, ok
    return res


def test_ebola_config(api):
    """configs[2]: Ebola Sierra Leone 2014 tree, times = seq(0, 1.36, length.out = 2001), N = 7e6,
    change points on the coarse grid, gamma ~ exp(N(3, 0.1)) (vignettes/Ebola_sierra_leone2014.Rmd:44-90)."""
    z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
    s = gen.summarize_phylo(z["edge"], z["edge_length"], int(z["ntip"]))
    times = np.linspace(0, 1.36, 2001)
    init = gen.coal_lik_init(s["samp_times"], s["n_sampled"], s["coal_times"], times[::50])
    cht = times[::50][1:-1]
    x_r, x_i = np.concatenate([[7e6], cht]), [len(cht), 2, 0, 1]
    rng = np.random.default_rng(1)
    B = 24
    param = np.empty((B, 42))
    param[:, 0] = 1.8 * np.exp(0.1 * rng.standard_normal(B))
    param[:, 1] = np.exp(3 + 0.1 * rng.standard_normal(B))
    param[:, 2:41] = np.exp(0.05 * rng.standard_normal((B, 39)))
    param[:, 41] = 20.0
    initial = np.tile([7e6, 20.0], (B, 1))
    eps = 0.3 * rng.standard_normal((B, 40, 2))
    res = check_batch(api, times, x_r, x_i, init, 1.36, param, initial, eps)
    assert np.all(np.isfinite(res["coalLog"]))


def test_synthetic_5000_tips(api):
    """configs[4]: volz_thin_sir genealogy with 5000 tips on a deterministic SIR trajectory (N = 1e6,
    R0 = 2.2, gamma = 0.2, I0 = 1, [0,100]), 2000 fine steps, gridsize 50."""
    times = np.linspace(0, 100, 2001)
    c = np.load(os.path.join(GOLDEN, "synthetic_5000.npz"))  # tests/golden/make_synthetic.py (8 min of thinning)
    init = gen.coal_lik_init(c["samp_times"], c["n_sampled"], c["coal_times"], times[::50])
    assert len(c["coal_times"]) == 4999 and len(init["C"]) > 5000
    cht = times[::50][1:-1]
    x_r, x_i = np.concatenate([[1e6], cht]), [len(cht), 2, 0, 1]
    rng = np.random.default_rng(2)
    B = 16
    param = np.empty((B, 42))
    param[:, 0] = 2.2 * np.exp(0.03 * rng.standard_normal(B))
    param[:, 1] = 0.2 * np.exp(0.03 * rng.standard_normal(B))
    param[:, 2:41] = np.exp(0.01 * rng.standard_normal((B, 39)))
    param[:, 41] = 20.0
    initial = np.tile([1e6, 1.0], (B, 1))
    eps = 0.1 * rng.standard_normal((B, 40, 2))
    eps[:, :12, :] = np.abs(eps[:, :12, :])  # keep the tiny early prevalence positive
    res = check_batch(api, times, x_r, x_i, init, 90.0, param, initial, eps)
    # the literal per-event kernel (E > 5000 events staged in shared memory) agrees with the per-cell path
    ll_ev = api.volz_loglik_nh2(init, res["LatentTraj"], res["betaN"], 90.0, [0, 1])
    good = res["coalLog"] != orc.SENTINEL
    assert np.max(np.abs(ll_ev[good] - res["coalLog"][good]) / np.abs(res["coalLog"][good])) < TOL


def test_seir_parameter_sweep(api, changepoint):
    """configs[3]: i.i.d. SEIR draws from log-normal priors, FULL evaluations only (p = 3, 3x3 Cholesky)."""
    g = changepoint
    rng = np.random.default_rng(3)
    B, x_i = 64, [39, 3, 0, 2]
    param = np.empty((B, 43))
    param[:, 0] = np.exp(0.8 + 0.15 * rng.standard_normal(B))
    param[:, 1] = np.exp(-0.7 + 0.1 * rng.standard_normal(B))
    param[:, 2] = np.exp(-1.6 + 0.1 * rng.standard_normal(B))
    param[:, 3:42] = np.exp(rng.standard_normal((B, 39)) / 30.0)
    param[:, 42] = np.exp(3 + 0.2 * rng.standard_normal(B))
    initial = np.tile([1e6, 4.0, 4.0], (B, 1))
    eps = 0.3 * rng.standard_normal((B, 40, 3))
    check_batch(api, g.times, g.x_r, x_i, g.init, g.t_correct, param, initial, eps, model="SEIR", min_ok=16)


def test_full_size_properties(api, changepoint):
    """At the bench's full batch size (4096 chains x 16 proposals) the oracle is too slow, so check
    properties that do not depend on size: determinism, permutation equivariance and batch-size
    invariance of every item (items are independent), plus spot parity of 32 items against the oracle."""
    g = changepoint
    B = 65536
    rng = np.random.default_rng(4)
    par = np.tile(g.final["par"].ravel(), (B, 1))
    par[:, 2] *= np.exp(0.05 * rng.standard_normal(B))
    par[:, 3] *= np.exp(0.05 * rng.standard_normal(B))
    par[:, 4:43] *= np.exp(0.02 * rng.standard_normal((B, 39)))
    th = rng.uniform(-0.3, 0.3, size=(B, 1, 1))
    eps = g.final["OriginTraj"][None] * np.cos(th) + rng.standard_normal((B, 40, 2)) * np.sin(th)
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    r1 = pb.full_loglik(par[:, 2:], par[:, :2], eps)
    r2 = pb.full_loglik(par[:, 2:], par[:, :2], eps)
    assert np.array_equal(r1["coalLog"], r2["coalLog"])                      # deterministic
    perm = rng.permutation(B)
    r3 = pb.full_loglik(par[perm, 2:], par[perm, :2], eps[perm])
    assert np.array_equal(r3["coalLog"], r1["coalLog"][perm])                # permutation equivariant
    sub = rng.choice(B, 1000, replace=False)
    r4 = pb.full_loglik(par[sub, 2:], par[sub, :2], eps[sub])               # small-batch path (other block size)
    assert np.array_equal(r4["coalLog"], r1["coalLog"][sub])
    cfg, ini = orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init)
    pick = sub[:32]
    ref, st = orc.full_loglik_batch(cfg, ini, par[pick, 2:], par[pick, :2], g.times, g.gridsize, eps[pick], g.t_correct)
    assert np.array_equal(st, r1["status"][pick])
    assert np.max(np.abs(ref - r1["coalLog"][pick]) / np.abs(ref)) < TOL
    assert 0 < np.mean(r1["coalLog"] == orc.SENTINEL) < 0.5


class _EbolaSetting:
    """The few attributes oracle/driver.py reads from a Golden, for the Ebola vignette's set-up."""

    def __init__(self):
        z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
        s = gen.summarize_phylo(z["edge"], z["edge_length"], int(z["ntip"]))
        self.coal = {k: s[k] for k in ("samp_times", "n_sampled", "coal_times")}
        self.times = np.linspace(0, 1.36, 2001)
        self.gridsize, self.t_correct = 50, 1.36
        self.init = gen.coal_lik_init(s["samp_times"], s["n_sampled"], s["coal_times"], self.times[::50])
        self.cht = self.times[::50][1:-1]
        self.x_r, self.x_i = np.concatenate([[7e6], self.cht]), [len(self.cht), 2, 0, 1]


def test_ebola_vignette_iteration_as_written():
    """vignettes/Ebola_sierra_leone2014.Rmd:76-88 AS WRITTEN: its Update_Param_each_Norm entry list is
    parIdlist = (1, 3, nch + 5), isLoglist = (1, 1, 0), priorIdlist = (1, 3, 5) -- a log random walk on S0 under pop_pr, one on
    R0 under gamma_pr, and the no-op hyper entry -- not the (gamma, hyper) pair of the other vignettes.  The driver runs it
    through the general entry list; chains must follow the restated R iteration (oracle/driver.py::eslice_iteration_general)
    step for step on the device-drawn streams, through both stages (burn1)."""
    from lnaphylodyn_b200 import api, mcmc, chains as chm, _lib
    from oracle import driver
    g = _EbolaSetting()
    nch = len(g.cht)
    prior = {"pop_pr": [1, 0.5, 1, 1], "R0_pr": [0.7, 0.5], "gamma_pr": [3, 0.1], "mu_pr": [3, 0.15], "hyper_pr": [3, 0.2]}
    proposal = {"pop_prop": 0.2, "R0_prop": 0.05, "gamma_prop": 0.09, "mu_prop": 0.1, "hyper_prop": 0.2}
    control = {"alpha": [1 / 7e6], "R0": 2.0, "gamma": 30.0, "ch": np.full(nch, 0.975), "traj": np.zeros((40, 2)), "hyper": 20.0}
    B, niter, burn1, seed = 4, 6, 3, 17
    res = mcmc.General_MCMC_with_ESlice(g.coal, g.times, 1.36, 7e6, gridsize=50, niter=niter, burn=0, thin=1,
                                        changetime=g.cht, DEMS=("S", "I"), prior=prior, proposal=proposal, control=control,
                                        ESS_vec=(0, 1, 0, 0, 1, 1, 0), likelihood="volz", model="SIR", Index=(0, 1), nparam=2,
                                        method="admcmc",
                                        options={"joint": True, "PCOV": None, "beta": 0.95, "burn1": burn1,
                                                 "parIdlist": {"a": 1, "c": 3, "d": nch + 5}, "isLoglist": {"a": 1, "c": 1, "d": 0},
                                                 "priorIdlist": {"a": 1, "c": 3, "d": 5}, "up": 1000000, "tune": 0.01, "hyper": False},
                                        n_chains=B, seed=seed)
    par0 = np.concatenate([[7e6, 1.0, 2.0, 30.0], control["ch"], [20.0]])
    st = [driver.init_state(g, par0, control["traj"]) for _ in range(B)]
    pl = [np.asarray(prior[k], dtype=float) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    base = {"priorList": [pl[0][:2], pl[1][:2], pl[2][:2], pl[4][:2]], "hyper": False,
            "entries": [(0, 1, (1.0, 0.5), 0.2), (2, 1, (3.0, 0.1), 0.09), (-1, 0, (3.0, 0.2), 0.2)]}
    pb = api.problem_for(g.times, g.x_r, g.x_i, 50, 1.36, g.init)
    n_norm = 43 + 80
    nrm, unf = np.empty((B, n_norm)), np.empty((B, chm.ITER_UNIF_GENERAL))
    for i in range(1, niter + 1):
        _lib.check(_lib.lib().lna_draw_streams(pb.h, B, 0, n_norm, chm.ITER_UNIF_GENERAL, seed, i, api._ptr(nrm), api._ptr(unf)))
        full = i >= burn1
        opts = dict(base, ESS_vec=[0, 1, 0, 0, 1 if full else 0, 1, 0], update_traj=full)
        for c in range(B):
            st[c] = driver.eslice_iteration_general(st[c], g, opts, nrm[c], unf[c])
            assert relerr(res["par"][c, i - 1], st[c]["par"]) < 1e-9, (i, c)
            assert res["l1"][c, i - 1] == pytest.approx(st[c]["coalLog"], rel=1e-9), (i, c)
    assert np.any(res["par"][:, -1, 0] != 7e6) and np.all(res["par"][:, :, 3] == 30.0)  # S0 walks, gamma is never updated
