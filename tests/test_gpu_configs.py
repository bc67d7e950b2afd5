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
