"""transX = "log": the reference's legacy log-scale LNA (SIR), SURVEY section 8(f.3) -- GPU against the C oracle (itself pinned
to the reference code as written for these variants, tests/test_ref_vs_oracle.py::test_integrator_and_lna_moments[SIR-log]).
Tolerance 1e-10 relative (north_star); sentinels and failures identical."""
import numpy as np
import pytest

from conftest import relerr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from lnaphylodyn_b200 import api as a
    return a


def setup(g, seed=3):
    rng = np.random.default_rng(seed)
    x_i = [39, 2, 0, 1]
    param = np.concatenate([[2.3, 0.21], np.exp(0.05 * rng.standard_normal(39)), [20.0]])
    initial = np.log([1e6, 10.0])
    return x_i, param, initial, g.times[:401]  # 8 LNA intervals: the log-scale Euler moments stay finite here


def test_log_scale_integrator_and_moments(api, changepoint):
    g = changepoint
    x_i, param, initial, t = setup(g)
    kw = dict(model="SIR", transX="log")
    ode_g, ode_o = api.ODE_rk45(initial, t, param, g.x_r, x_i, **kw), orc.ODE_rk45(initial, t, param, g.x_r, x_i, **kw)
    assert relerr(ode_g[:, 1:], ode_o[:, 1:]) < TOL
    tol = np.log(50.0)
    assert api.ODE_rk45_stop(initial, t, param, g.x_r, x_i, tol=tol, **kw) == orc.ODE_rk45_stop(initial, t, param, g.x_r, x_i, tol=tol, **kw)
    seg = ode_o[100:150]
    a, b = api.SigmaF(seg, param, g.x_r, x_i, **kw), orc.SigmaF(seg, param, g.x_r, x_i, **kw)
    assert relerr(a["expF"], b["expF"]) < TOL and relerr(a["Sigma"], b["Sigma"]) < 1e-9
    for fn in ("KF_param", "KF_param_chol"):
        a, b = getattr(api, fn)(ode_o, param, 50, g.x_r, x_i, **kw), getattr(orc, fn)(ode_o, param, 50, g.x_r, x_i, **kw)
        assert relerr(a["A"], b["A"]) < TOL
        assert np.max(np.abs(a["Sigma"] - b["Sigma"])) < 1e-9 * np.max(np.abs(b["Sigma"]))
    a = api.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    b = orc.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    assert relerr(a["FT"]["A"], b["FT"]["A"]) < TOL
    assert np.max(np.abs(a["FT"]["Sigma"] - b["FT"]["Sigma"])) < 1e-9 * np.max(np.abs(b["FT"]["Sigma"]))
    assert relerr(a["Ode"][:, 1:], b["Ode"][:, 1:]) < TOL and relerr(a["betaN"], b["betaN"]) < 1e-14


def test_log_scale_full_evaluation_and_likelihoods(api, changepoint):
    """New_Param_List -> TransformTraj -> volz_loglik_nh2 on the log scale as ONE batched call, against the oracle's composition;
    and the likelihood functions on a log-scale trajectory (volz_loglik_nh2/nh3 "log", coal_loglik3 "log")."""
    g = changepoint
    x_i, param, initial, _ = setup(g)
    rng = np.random.default_rng(8)
    B = 24
    P = np.tile(param, (B, 1))
    P[:, 0] *= np.exp(0.05 * rng.standard_normal(B))
    P[:, 1] *= np.exp(0.05 * rng.standard_normal(B))
    eps = 0.3 * rng.standard_normal((B, 40, 2))
    I0 = np.tile(initial, (B, 1))
    pb = api.problem_for(g.times, g.x_r, x_i, g.gridsize, g.t_correct, g.init, "SIR", "log")
    res = pb.full_loglik(P, I0, eps, want=("coalLog", "LatentTraj", "betaN"))
    n_ok = 0
    for b in range(B):
        try:
            npl = orc.New_Param_List(P[b], initial, g.gridsize, g.times, g.x_r, x_i, model="SIR", transX="log")
        except ValueError:
            assert res["status"][b] & 1 and res["coalLog"][b] == orc.SENTINEL
            continue
        lat = orc.TransformTraj(npl["Ode"], eps[b], npl["FT"])
        want = orc.volz_loglik_nh2(g.init, lat, npl["betaN"], g.t_correct, [0, 1], "log")
        if want == orc.SENTINEL:
            assert res["coalLog"][b] == orc.SENTINEL
            continue
        n_ok += 1
        assert res["coalLog"][b] == pytest.approx(want, rel=1e-9)
        assert np.max(np.abs(res["LatentTraj"][b][:, 1:] - lat[:, 1:])) < 1e-9 * max(1.0, np.max(np.abs(lat[:, 1:])))
    assert n_ok >= B // 2
    # likelihood functions on a given log-scale trajectory
    lt = np.column_stack([g.final["LatentTraj"][:, 0], np.log(np.abs(g.final["LatentTraj"][:, 1:]) + 1.0)])
    bn = g.final["betaN"].ravel()
    for fn in ("volz_loglik_nh2", "volz_loglik_nh3"):
        assert getattr(api, fn)(g.init, lt, bn, g.t_correct, [0, 1], "log") == pytest.approx(
            getattr(orc, fn)(g.init, lt, bn, g.t_correct, [0, 1], "log"), rel=TOL)
    assert api.coal_loglik3(g.init, lt, g.t_correct, 30.0, 1, "log") == pytest.approx(
        orc.coal_loglik3(g.init, lt, g.t_correct, 30.0, 1, "log"), rel=TOL)


def test_log_scale_is_refused_where_the_reference_never_uses_it(api, changepoint):
    """SEIR on the log scale and the Metropolis / slice operators on a log-scale problem fail loudly (no fallback)."""
    from lnaphylodyn_b200._lib import LnaError
    g = changepoint
    with pytest.raises(LnaError, match="transX"):
        api.Problem(g.times, g.x_r, [39, 3, 0, 2], g.gridsize, g.t_correct, g.init, model="SEIR", transX="log")
    par = g.final["par"].ravel()
    rng = np.random.default_rng(1)
    with pytest.raises(LnaError, match="standard"):
        api.ESlice_par_General(par, g.times, g.final["OriginTraj"], [(1, 1), (0.7, 0.5), (-1.7, 0.1), (3, 0.2)], g.x_r, g.x_i,
                               g.init, g.gridsize, [0, 1, 0, 0, 1, 1, 0], -2600.0, g.t_correct, transX="log",
                               normals=rng.standard_normal(43), uniforms=rng.uniform(size=22))
