"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): log-likelihoods and trajectories within 1e-10 relative, FP64.
"""
import numpy as np
import pytest

from conftest import relerr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    import lnaphylodyn_b200 as m
    from lnaphylodyn_b200 import _lib
    assert _lib.lib().lna_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return m.api


def draw_inputs(g, B, seed, scale=0.05):
    """Parameter draws around the cached final state + fresh N(0,1) latent increments."""
    rng = np.random.default_rng(seed)
    par = np.tile(g.final["par"].ravel(), (B, 1))
    par[:, 2] *= np.exp(scale * rng.standard_normal(B))            # R0
    par[:, 3] *= np.exp(scale * rng.standard_normal(B))            # gamma
    par[:, 4:43] *= np.exp(0.5 * scale * rng.standard_normal((B, 39)))  # change ratios
    eps = rng.standard_normal((B, 40, 2))
    # half of the draws stay close to the cached latent path (valid trajectories), half are wild
    th = rng.uniform(-0.3, 0.3, size=(B // 2, 1, 1))
    eps[: B // 2] = g.final["OriginTraj"][None] * np.cos(th) + eps[: B // 2] * np.sin(th)
    return par, eps


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_full_loglik_final_state_vs_golden(api, name, request):
    g = request.getfixturevalue(name)
    par = g.final["par"].ravel()
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    res = pb.full_loglik(par[2:], par[:2], g.final["OriginTraj"], want=("coalLog", "FT", "Ode", "betaN", "LatentTraj"))
    assert res["status"] == 0
    assert relerr(res["Ode"], g.final["Ode_Traj_coarse"]) < TOL
    assert relerr(res["FT"]["A"], g.final["FT.A"]) < TOL
    assert np.max(np.abs(res["FT"]["Sigma"] - g.final["FT.Sigma"])) < 1e-9
    assert relerr(res["betaN"], g.final["betaN"].ravel()) < TOL
    assert relerr(res["LatentTraj"][:, 1:], g.final["LatentTraj"][:, 1:]) < TOL
    assert abs(res["coalLog"] - float(g.final["coalLog"][0])) < TOL * abs(res["coalLog"])


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_full_loglik_batch_vs_oracle(api, name, request):
    g = request.getfixturevalue(name)
    B = 96
    par, eps = draw_inputs(g, B, seed=11)
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    res = pb.full_loglik(par[:, 2:], par[:, :2], eps, want=("coalLog", "FT", "Ode", "betaN", "LatentTraj"))
    res_ll_only = pb.full_loglik(par[:, 2:], par[:, :2], eps)
    cfg, ini = orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init)
    worst, worst_ft = 0.0, {"A": 0.0, "Sigma": 0.0, "LatentTraj": 0.0}
    for b in range(B):
        o = orc.full_loglik(cfg, ini, par[b, 2:], par[b, :2], g.times, g.gridsize, eps[b], g.t_correct)
        assert o["status"] == res["status"][b]
        if o["coalLog"] == orc.SENTINEL:
            assert res["coalLog"][b] == orc.SENTINEL
        else:
            worst = max(worst, abs(res["coalLog"][b] - o["coalLog"]) / abs(o["coalLog"]))
        assert res_ll_only["coalLog"][b] == res["coalLog"][b]
        assert relerr(res["Ode"][b], o["Ode"]) < TOL
        assert relerr(res["betaN"][b], o["betaN"]) < TOL
        # FT$A, FT$Sigma (= L) and the latent states: 1e-10 relative to the size of the object each entry belongs to --
        # the 2 x 2 matrix of its LNA interval, the trajectory column.  Entry-wise relative error is not a meaningful bar
        # for the entries that are differences of O(1) quantities (A's off-diagonals of 1e-6 next to diagonals of 1, the
        # Cholesky factor's l11 = sqrt(s11 - l10^2), a latent I of a few individuals next to S of 1e6).
        for key in ("A", "Sigma"):
            d, ref = res["FT"][key][b], o["FT"][key]
            scale = np.max(np.abs(ref), axis=(0, 1), keepdims=True)
            worst_ft[key] = max(worst_ft[key], float(np.max(np.abs(d - ref) / scale)))
        lat, lref = res["LatentTraj"][b][:, 1:], o["LatentTraj"][:, 1:]
        worst_ft["LatentTraj"] = max(worst_ft["LatentTraj"], float(np.max(np.abs(lat - lref) / np.max(np.abs(lref), axis=0))))
    print("worst scaled errors:", worst_ft, "coalLog", worst)
    assert all(v < TOL for v in worst_ft.values()), worst_ft
    assert worst < TOL, worst
    assert np.sum(res["coalLog"] != orc.SENTINEL) >= B // 4 and np.sum(res["coalLog"] == orc.SENTINEL) >= 1


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_per_iteration_known_answers_event_kernel(api, name, request):
    """(par_i, Trajectory_i) -> coalLog_i of the reference's own MCMC run, via the per-event kernel."""
    g = request.getfixturevalue(name)
    traj = g.iters["Trajectory"]
    betaN = np.stack([orc.betaTs(par[2:], tr[:, 0], g.x_r, g.x_i) for par, tr in zip(g.iters["par"], traj)])
    ll = api.volz_loglik_nh2(g.init, traj, betaN, g.t_correct, g.x_i[2:4])
    assert np.max(np.abs(ll - g.iters["coalLog"]) / np.abs(g.iters["coalLog"])) < TOL
    b1 = api.betaTs(g.iters["par"][5][2:], traj[5][:, 0], g.x_r, g.x_i)
    assert relerr(b1, betaN[5]) < 1e-14


def test_cell_sums_match_events(api, changepoint):
    g = changepoint
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    cd, yy, cnt = pb.cell_sums()
    start = np.concatenate([[0], np.cumsum(g.init["gridrep"][:-1])]).astype(int)
    for i in range(len(cd)):
        sl = slice(start[i], start[i + 1])
        assert cnt[i] == g.init["gridrep"][i]
        assert cd[i] == pytest.approx(np.sum(g.init["C"][sl] * g.init["D"][sl]), rel=1e-13, abs=1e-300)
        assert yy[i] == np.sum(g.init["y"][sl])


def test_negative_state_sentinel(api, changepoint):
    """A huge negative latent increment drives I < 0 inside the checked rows -> -1e7 (A.9)."""
    g = changepoint
    par = g.final["par"].ravel()
    eps = np.zeros((40, 2))
    eps[3, 1] = -1e6
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    res = pb.full_loglik(par[2:], par[:2], eps)
    cfg, ini = orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init)
    o = orc.full_loglik(cfg, ini, par[2:], par[:2], g.times, g.gridsize, eps, g.t_correct)
    assert o["coalLog"] == orc.SENTINEL and res["coalLog"] == orc.SENTINEL
    assert res["status"] & 2


def test_cholesky_failure_flag(api, changepoint):
    """NaN parameters make arma::chol throw in the reference; here: status bit + -1e7, and the
    Rcpp-style wrapper raises like KF_param_chol (LNA_functional.cpp:660-664)."""
    g = changepoint
    par = g.final["par"].ravel().copy()
    par[2] = np.nan
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    res = pb.full_loglik(par[2:], par[:2], g.final["OriginTraj"])
    assert res["status"] & 1 and res["coalLog"] == orc.SENTINEL
    with pytest.raises(ValueError):
        api.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
    with pytest.raises(ValueError):
        orc.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)


def test_rcpp_mirrors_vs_oracle(api, changepoint):
    g = changepoint
    par = g.final["par"].ravel()
    param, initial = par[2:], par[:2]
    o_ode = orc.ODE_rk45(initial, g.times, param, g.x_r, g.x_i)
    d_ode = api.ODE_rk45(initial, g.times, param, g.x_r, g.x_i)
    assert relerr(d_ode, np.where(o_ode == 0, 0, o_ode)) < TOL or np.max(np.abs(d_ode - o_ode) / np.maximum(np.abs(o_ode), 1e-300)[...]) < TOL
    for fn in ("KF_param", "KF_param_chol"):
        o = getattr(orc, fn)(o_ode, param, g.gridsize, g.x_r, g.x_i)
        d = getattr(api, fn)(o_ode, param, g.gridsize, g.x_r, g.x_i)
        assert relerr(d["A"], o["A"]) < TOL
        assert np.max(np.abs(d["Sigma"] - o["Sigma"])) < 1e-9 * max(1.0, np.max(np.abs(o["Sigma"])))
    assert np.array_equal(api.Ode_Coarse_Slicer(o_ode, g.gridsize), orc.Ode_Coarse_Slicer(o_ode, g.gridsize))
    npl_o = orc.New_Param_List(param, initial, g.gridsize, g.times, g.x_r, g.x_i)
    npl_d = api.New_Param_List(param, initial, g.gridsize, g.times, g.x_r, g.x_i)
    assert relerr(npl_d["Ode"], npl_o["Ode"]) < TOL and relerr(npl_d["betaN"], npl_o["betaN"]) < TOL
    lat_o = orc.TransformTraj(npl_o["Ode"], g.final["OriginTraj"], npl_o["FT"])
    lat_d = api.TransformTraj(npl_o["Ode"], g.final["OriginTraj"], npl_o["FT"])
    assert relerr(lat_d[:, 1:], lat_o[:, 1:]) < TOL
    sf_o = orc.SigmaF(o_ode[100:150], param, g.x_r, g.x_i)
    sf_d = api.SigmaF(o_ode[100:150], param, g.x_r, g.x_i)
    assert relerr(sf_d["expF"], sf_o["expF"]) < TOL and relerr(sf_d["Sigma"], sf_o["Sigma"]) < 1e-9
    pt_o = orc.param_transform(33.3, param, g.x_r, g.x_i)
    pt_d = api.param_transform(33.3, param, g.x_r, g.x_i)
    assert relerr(pt_d, pt_o) < 1e-14
    rng = np.random.default_rng(3)
    nrm = rng.standard_normal(80)
    s_o = orc.Traj_sim_general_noncentral(npl_o["Ode"], npl_o["FT"], g.t_correct, nrm)
    s_d = api.Traj_sim_general_noncentral(npl_o["Ode"], npl_o["FT"], g.t_correct, normals=nrm)
    assert relerr(s_d["SimuTraj"][:, 1:], s_o["SimuTraj"][:, 1:]) < TOL
    assert np.array_equal(s_d["OriginTraj"], s_o["OriginTraj"])
    assert s_d["logOrigin"] == pytest.approx(s_o["logOrigin"], rel=1e-13)
    assert s_d["logMultiNorm"] == pytest.approx(s_o["logMultiNorm"], rel=1e-10)
    k_o = orc.coal_loglik3(g.init, lat_o, g.t_correct, 3.0, 1)
    k_d = api.coal_loglik3(g.init, lat_o, g.t_correct, 3.0, 1)
    assert k_d == pytest.approx(k_o, rel=TOL)
    logtraj = lat_o.copy()
    logtraj[:, 1:] = np.log(lat_o[:, 1:])
    assert api.coal_loglik(g.init, logtraj, g.t_correct, 3.0) == pytest.approx(orc.coal_loglik(g.init, logtraj, g.t_correct, 3.0), rel=TOL)


def test_minor_exports(api, changepoint):
    g = changepoint
    par = g.final["par"].ravel()
    param, initial = par[2:], par[:2]
    lat = g.final["LatentTraj"]
    bn = g.final["betaN"].ravel()
    assert api.volz_loglik_nh3(g.init, lat, bn, g.t_correct, [0, 1]) == pytest.approx(
        orc.volz_loglik_nh3(g.init, lat, bn, g.t_correct, [0, 1]), rel=TOL)
    bad = lat.copy()
    bad[40, 2] = -1.0  # beyond the checked rows: log(negative) -> NaN; nh2 guards it, nh3 does not
    assert api.volz_loglik_nh2(g.init, bad, bn, 100.0, [0, 1]) == orc.volz_loglik_nh2(g.init, bad, bn, 100.0, [0, 1]) == orc.SENTINEL
    assert np.isnan(api.volz_loglik_nh3(g.init, bad, bn, 100.0, [0, 1])) and np.isnan(orc.volz_loglik_nh3(g.init, bad, bn, 100.0, [0, 1]))
    assert api.log_like_traj_general_adjust(lat, lat, None, 50, 90) == 0.0
    for tol in (5.0, 2000.0):
        assert api.ODE_rk45_stop(initial, g.times, param, g.x_r, g.x_i, tol=tol) == orc.ODE_rk45_stop(initial, g.times, param, g.x_r, g.x_i, tol=tol)
    nrm = np.random.default_rng(8).standard_normal(80)
    s_d = api.Traj_sim_ezG_NC(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=nrm)
    npl = orc.New_Param_List(param, initial, 50, g.times, g.x_r, g.x_i)
    s_o = orc.Traj_sim_general_noncentral(npl["Ode"], npl["FT"], g.t_correct, nrm)
    assert relerr(s_d["SimuTraj"][:, 1:], s_o["SimuTraj"][:, 1:]) < TOL


def test_update_param_decisions(api, changepoint):
    g = changepoint
    B = 64
    par, _ = draw_inputs(g, B, seed=5, scale=0.01)
    eps = g.final["OriginTraj"]
    rng = np.random.default_rng(9)
    u = rng.uniform(size=B)
    coal_log = float(g.final["coalLog"][0])
    d = api.Update_Param(par[:, 2:], par[:, :2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, coal_log, 0.1,
                         g.t_correct, unif=u)
    flips = 0
    for b in range(B):
        o = orc.Update_Param(par[b, 2:], par[b, :2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, coal_log, 0.1,
                             g.t_correct, unif=u[b])
        flips += int(bool(o["accept"]) != bool(d["accept"][b]))
        ref = o["coalLog"] if o["accept"] else o["coalLog_proposed"]
        assert d["coalLog"][b] == pytest.approx(ref, rel=TOL)
    assert flips == 0
    assert 0 < d["accept"].sum() < B  # the test exercises both branches


def test_seir_full_loglik_vs_oracle(api, changepoint):
    """p = 3 model variant (model="SEIR": states S,E,I; params R0, mu, gamma; volz index (0,2))."""
    g = changepoint
    B = 48
    rng = np.random.default_rng(21)
    x_i = [39, 3, 0, 2]
    param = np.empty((B, 43))
    param[:, 0] = 2.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 1] = 0.5 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 2] = 0.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 3:42] = np.exp(0.02 * rng.standard_normal((B, 39)))
    param[:, 42] = 20.0
    initial = np.tile([1e6, 5.0, 5.0], (B, 1))
    eps = 0.5 * rng.standard_normal((B, 40, 3))
    pb = api.problem_for(g.times, g.x_r, x_i, g.gridsize, g.t_correct, g.init, model="SEIR")
    res = pb.full_loglik(param, initial, eps, want=("coalLog", "FT", "Ode", "betaN", "LatentTraj"))
    cfg, ini = orc.Cfg(g.x_r, x_i, 43, model="SEIR"), orc.Init(g.init)
    n_ok = 0
    for b in range(B):
        o = orc.full_loglik(cfg, ini, param[b], initial[b], g.times, g.gridsize, eps[b], g.t_correct)
        assert o["status"] == res["status"][b]
        assert relerr(res["Ode"][b], o["Ode"]) < TOL
        assert relerr(res["FT"]["A"][b], o["FT"]["A"]) < 1e-9
        assert np.max(np.abs(res["FT"]["Sigma"][b] - o["FT"]["Sigma"])) < 1e-8 * max(1.0, np.max(np.abs(o["FT"]["Sigma"])))
        if o["coalLog"] == orc.SENTINEL:
            assert res["coalLog"][b] == orc.SENTINEL
        else:
            n_ok += 1
            assert res["coalLog"][b] == pytest.approx(o["coalLog"], rel=TOL)
    assert n_ok > B // 2


def test_centred_entry_points(api, changepoint):
    """KF_param / log_like_traj_general2 / Traj_sim_general3 / Traj_sim_ezG2 / ESlice_general2 and the
    reference's own 'Test SIR model' identity (testStructureCoal.R:41-51) through the GPU path."""
    g = changepoint
    par = g.final["par"].ravel()
    param, initial = par[2:], par[:2]
    ode = orc.ODE_rk45(initial, g.times, param, g.x_r, g.x_i)
    kf = orc.KF_param(ode, param, 50, g.x_r, g.x_i)
    kfc = orc.KF_param_chol(ode, param, 50, g.x_r, g.x_i)
    co = orc.Ode_Coarse_Slicer(ode, 50)
    rng = np.random.default_rng(12)
    z = rng.standard_normal(80)
    s_o = orc.Traj_sim_general3(co, kf, g.t_correct, normals=z)
    s_d = api.Traj_sim_general3(co, kf, g.t_correct, normals=z)
    assert relerr(s_d["SimuTraj"][:, 1:], s_o["SimuTraj"][:, 1:]) < TOL and s_d["SimuTraj"][0, 0] == 0.0
    assert s_d["loglike"] == pytest.approx(s_o["loglike"], rel=1e-9)
    e_d = api.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
    e_o = orc.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
    assert relerr(e_d["SimuTraj"][:, 1:], e_o["SimuTraj"][:, 1:]) < 1e-9 and e_d["loglike"] == pytest.approx(e_o["loglike"], rel=1e-8)
    l_o = orc.log_like_traj_general2(s_o["SimuTraj"], co, kf, 50, g.t_correct)
    l_d = api.log_like_traj_general2(s_o["SimuTraj"], co, kf, 50, g.t_correct)
    assert l_d == pytest.approx(l_o, rel=1e-9)
    # testthat identity (c): non-centred simulation density == centred density (up to the 1e-8 jitter)
    nc = api.Traj_sim_general_noncentral(co, kfc, 150.0, normals=z)
    assert nc["logMultiNorm"] + nc["logOrigin"] == pytest.approx(
        api.log_like_traj_general2(nc["SimuTraj"], co, kf, 50, 150.0), rel=1.5e-8)
    # centred elliptical slice, step for step
    bn = orc.betaTs(param, co[:, 0], g.x_r, g.x_i)
    B = 24
    nrm, unf = rng.standard_normal((B, 80)), rng.random((B, 22))
    unf[:3, 0] = 1.0 - 1e-12
    rep = lambda a: np.tile(a, (B,) + (1,) * np.ndim(a))
    d = api.ESlice_general2(rep(g.final["LatentTraj"]), rep(co), {"A": rep(kf["A"]), "Sigma": rep(kf["Sigma"])}, initial,
                            g.init, rep(bn), g.t_correct, volz=True, normals=nrm, uniforms=unf)
    for b in range(B):
        o = orc.ESlice_general2(g.final["LatentTraj"], co, kf, initial, g.init, bn, g.t_correct, volz=True,
                                normals=nrm[b], uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"] and (d["status"][b] & 8) == (o["status"] & 8), b
        assert relerr(d["newTraj"][b][:, 1:], o["newTraj"][:, 1:]) < 1e-9
    assert d["n_evals"].max() > 3


def test_seasonal_sirs_model_variant(api):
    """src/SIRS_period.cpp as a model variant of the ODE / LNA-moment kernels (configs[3]): ODE2,
    SIRS_KOM_Filter and the pseudo-inverse trajectory density, 64 parameter draws."""
    t = np.linspace(0, 100, 2001)
    rng = np.random.default_rng(31)
    B = 64
    param = np.column_stack([3e-6 * np.exp(0.1 * rng.standard_normal(B)), 0.2 * np.exp(0.1 * rng.standard_normal(B)),
                             0.02 * np.exp(0.1 * rng.standard_normal(B)), rng.uniform(0.1, 0.5, B)])
    init = np.tile([9e4, 100.0, 9900.0], (B, 1))
    d_ode = api.ODE2(init, t, param, 40.0)
    for b in range(0, B, 8):
        o_ode = orc.ODE2(init[b], t, param[b], 40.0)
        assert np.max(np.abs(d_ode[b] - o_ode) / np.maximum(np.abs(o_ode), 1e-300)[...]) < TOL
        kf_o = orc.SIRS_KOM_Filter(o_ode, param[b], 50, 40.0)
        kf_d = api.SIRS_KOM_Filter(o_ode, param[b], 50, 40.0)
        assert relerr(kf_d["A"], kf_o["A"]) < 1e-9
        assert np.max(np.abs(kf_d["Sigma"] - kf_o["Sigma"])) < 1e-9 * np.max(np.abs(kf_o["Sigma"]))
        co = orc.Ode_Coarse_Slicer(o_ode, 50)
        dev = rng.standard_normal((41, 3)) * 30.0
        dev[:, 2] = -dev[:, 0] - dev[:, 1]           # stay on the conserved-population plane
        dev[0] = 0
        sde = co.copy()
        sde[:, 1:] += dev
        l_o = orc.log_like_trajSIRS(sde, co, kf_o, 50, 90.0)
        l_d = api.log_like_trajSIRS(sde, co, kf_o, 50, 90.0)
        assert np.isfinite(l_o) and l_d == pytest.approx(l_o, rel=1e-7)


@pytest.mark.parametrize("n_chains,proposals", [(1, 1), (3, 16), (700, 7), (4096, 16)])
def test_chain_shared_full_loglik_equals_per_item(api, changepoint, n_chains, proposals):
    """lna_full_loglik_shared (one initial state and one OriginTraj per chain, `proposals` parameter vectors each) is the
    per-item call on the expanded inputs, bit for bit -- synchronous (chunked above 32k evaluations), asynchronous, and for
    odd shapes; spot-checked against the oracle."""
    g = changepoint
    rng = np.random.default_rng(n_chains)
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    par0 = g.final["par"].ravel()
    param = np.tile(par0[2:], (n_chains, proposals, 1))
    param[:, :, 0] *= np.exp(0.05 * rng.standard_normal((n_chains, proposals)))
    param[:, :, 1] *= np.exp(0.05 * rng.standard_normal((n_chains, proposals)))
    param[:, :, 2:41] *= np.exp(0.02 * rng.standard_normal((n_chains, proposals, 39)))
    initial = np.tile(par0[:2], (n_chains, 1)) * np.exp(0.01 * rng.standard_normal((n_chains, 2)))
    eps = g.final["OriginTraj"][None] + 0.1 * rng.standard_normal((n_chains, 40, 2))
    sh = pb.full_loglik_shared(param, initial, eps)
    flat = pb.full_loglik(param.reshape(-1, 42), np.repeat(initial, proposals, axis=0), np.repeat(eps, proposals, axis=0))
    assert np.array_equal(sh["coalLog"].ravel(), np.atleast_1d(flat["coalLog"]))
    assert np.array_equal(sh["status"].ravel(), np.atleast_1d(flat["status"]))
    sh2 = pb.full_loglik_shared(param, initial, eps, asynchronous=True)
    assert np.array_equal(sh2["coalLog"], sh["coalLog"]) and np.array_equal(sh2["status"], sh["status"])
    cfg, ini = orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init)
    for c in sorted({0, n_chains // 2, n_chains - 1}):
        ref, _ = orc.full_loglik_batch(cfg, ini, param[c], np.tile(initial[c], (proposals, 1)), g.times, g.gridsize,
                                       np.tile(eps[c], (proposals, 1, 1)), g.t_correct)
        assert relerr(sh["coalLog"][c], ref) < TOL


def _sirs_draws(rng, B):
    return np.column_stack([3e-6 * np.exp(0.1 * rng.standard_normal(B)), 0.2 * np.exp(0.1 * rng.standard_normal(B)),
                            0.02 * np.exp(0.1 * rng.standard_normal(B)), rng.uniform(0.1, 0.5, B)])


def _sirs_data(t, rng, truth=(3e-6, 0.2, 0.02, 0.3), init=(9e4, 100.0, 9900.0), scale=30.0):
    """A 'data' trajectory for the sweep: the ODE at the true parameters plus a perturbation on the conserved-population plane."""
    co = orc.Ode_Coarse_Slicer(orc.ODE2(np.array(init), t, np.array(truth), 40.0), 50)
    dev = rng.standard_normal((41, 3)) * scale
    dev[:, 2] = -dev[:, 0] - dev[:, 1]
    dev[0] = 0
    sde = co.copy()
    sde[:, 1:] += dev
    return sde


def _sirs_oracle(init, t, param, sde, t_correct=90.0):
    ode = orc.ODE2(np.asarray(init, dtype=float), t, param, 40.0)
    kf = orc.SIRS_KOM_Filter(ode, param, 50, 40.0)
    co = orc.Ode_Coarse_Slicer(ode, 50)
    return orc.log_like_trajSIRS(sde, co, kf, 50, t_correct), kf, co


def test_seasonal_sirs_fused_sweep_vs_oracle(api):
    """The fused seasonal-SIRS evaluation (one thread per draw: ODE2 -> SIRS_KOM_Filter -> log_like_trajSIRS in registers)
    against the oracle's composition of the three exports, draw by draw: moments at 1e-10 relative to the largest entry of
    each interval's matrix, coarse ODE at 1e-10.  The density goes through the closed-form pseudo-inverse of a covariance
    whose entries cancel to its rank-2 part (m11 - m01 - lambda in the eigenvectors): the reference's own formula turns a
    1e-14 relative change of Sigma into up to ~1e-9 of the log-density on these draws, so the log-density is held to 1e-8
    relative here and to 1e-12 when BOTH sides are given the same moments (test_seasonal_sirs_model_variant)."""
    t = np.linspace(0, 100, 2001)
    rng = np.random.default_rng(32)
    B = 48
    param = _sirs_draws(rng, B)
    init = np.array([9e4, 100.0, 9900.0])
    sde = _sirs_data(t, rng)
    d = api.SIRS_loglik_sweep(param, init, t, sde, 50, 90.0, 40.0, want=("loglik", "A", "Sigma", "Ode"))
    assert not d["status"].any()
    worst = 0.0
    for b in range(B):
        l_o, kf, co = _sirs_oracle(init, t, param[b], sde)
        assert relerr(d["Ode"][b][:, 1:], co[:, 1:]) < TOL and np.array_equal(d["Ode"][b][:, 0], co[:, 0])
        for i in range(40):
            assert np.max(np.abs(d["A"][b][:, :, i] - kf["A"][:, :, i])) < TOL * np.max(np.abs(kf["A"][:, :, i]))
            assert np.max(np.abs(d["Sigma"][b][:, :, i] - kf["Sigma"][:, :, i])) < TOL * np.max(np.abs(kf["Sigma"][:, :, i]))
        assert np.isfinite(l_o)
        worst = max(worst, abs(d["loglik"][b] - l_o) / abs(l_o))
    assert worst < 1e-8, worst
    # per-draw initial states and per-draw / per-group trajectories are the same evaluation
    d2 = api.SIRS_loglik_sweep(param, np.tile(init, (B, 1)), t, np.tile(sde, (B, 1, 1)), 50, 90.0, 40.0)
    d3 = api.SIRS_loglik_sweep(param, init, t, np.tile(sde, (B // 8, 1, 1)), 50, 90.0, 40.0)
    assert np.array_equal(d2["loglik"], d["loglik"]) and np.array_equal(d3["loglik"], d["loglik"])
    # rows later than t_correct do not count (:104)
    early = api.SIRS_loglik_sweep(param, init, t, sde, 50, 50.0, 40.0)["loglik"]
    l_e, _, _ = _sirs_oracle(init, t, param[0], sde, 50.0)
    assert early[0] == pytest.approx(l_e, rel=1e-8) and not np.allclose(early, d["loglik"])


def test_seasonal_sirs_fused_sweep_full_size(api):
    """BASELINE configs[3]: 16 384 draws per sweep in ONE launch.  Size-independent properties: the sweep equals the same
    draws evaluated in small batches (bit for bit: a draw's result does not depend on its neighbours or its position),
    permuting the draws permutes the results, duplicated draws give duplicated results; spot-checked against the oracle."""
    t = np.linspace(0, 100, 2001)
    rng = np.random.default_rng(33)
    B = 16384
    param = _sirs_draws(rng, B)
    param[1000] = param[7]
    init = np.array([9e4, 100.0, 9900.0])
    sde = _sirs_data(t, rng)
    full = api.SIRS_loglik_sweep(param, init, t, sde, 50, 90.0, 40.0)["loglik"]
    assert np.all(np.isfinite(full)) and full[1000] == full[7]
    perm = rng.permutation(B)
    assert np.array_equal(api.SIRS_loglik_sweep(param[perm], init, t, sde, 50, 90.0, 40.0)["loglik"], full[perm])
    for lo in (0, 4096 - 13, 16384 - 100):
        part = api.SIRS_loglik_sweep(param[lo:lo + 100], init, t, sde, 50, 90.0, 40.0)["loglik"]
        assert np.array_equal(part, full[lo:lo + 100])
    for b in (0, 1, 5000, 12345, 16383):
        l_o, _, _ = _sirs_oracle(init, t, param[b], sde)
        assert full[b] == pytest.approx(l_o, rel=1e-8)
    # the likelihood surface is informative: the sweep ranks draws near the truth above draws far from it
    near = np.argsort(np.abs(np.log(param[:, 0] / 3e-6)) + np.abs(np.log(param[:, 1] / 0.2)))
    assert full[near[:200]].mean() > full[near[-200:]].mean()


def test_traj_sim_sirs_reports_cholesky_outcome_per_item(api):
    """Traj_sim_SIRS (SIRS_period.cpp:197-240): noise = chol(Sigma + 1e-13 I)' z of a RANK-2 covariance -- whether the
    factorisation exists hangs on the last rounding of Sigma, so the outcome is reported per item and compared with the
    oracle (which raises where the reference's arma::chol throws): same verdict on the same (A, Sigma), same trajectory and
    density where it succeeds; a covariance lifted off the singular plane always factorises."""
    t = np.linspace(0, 100, 2001)
    rng = np.random.default_rng(41)
    B = 40
    param = _sirs_draws(rng, B)
    init = np.array([9e4, 100.0, 9900.0])
    ok_o = ok_d = both = 0
    for b in range(B):
        ode = orc.ODE2(init, t, param[b], 40.0)
        kf = orc.SIRS_KOM_Filter(ode, param[b], 50, 40.0)
        co = orc.Ode_Coarse_Slicer(ode, 50)
        z = rng.standard_normal(120)
        try:
            o = orc.Traj_sim_SIRS(init, co, kf, 90.0, normals=z)
        except ValueError:
            o = None
        try:
            d = api.Traj_sim_SIRS(init, co, kf, 90.0, normals=z)
        except ValueError:
            d = None
        ok_o += o is not None
        ok_d += d is not None
        assert (o is None) == (d is None), b   # same arithmetic on the same Sigma: same verdict
        if o is not None and d is not None:
            both += 1
            assert relerr(d["SimuTraj"][:, 1:], o["SimuTraj"][:, 1:]) < 1e-9
            assert d["loglike"] == pytest.approx(o["loglike"], rel=1e-7)
        # well-conditioned variant: Sigma + 1e-3 I always factorises, on both sides
        kf2 = {"A": kf["A"], "Sigma": kf["Sigma"] + 1e-3 * np.eye(3)[:, :, None]}
        o2, d2 = orc.Traj_sim_SIRS(init, co, kf2, 90.0, normals=z), api.Traj_sim_SIRS(init, co, kf2, 90.0, normals=z)
        assert relerr(d2["SimuTraj"][:, 1:], o2["SimuTraj"][:, 1:]) < TOL
    # batch form: the per-item status says which draws the reference would have thrown on
    ode = orc.ODE2(init, t, param[0], 40.0)
    kf, co = orc.SIRS_KOM_Filter(ode, param[0], 50, 40.0), orc.Ode_Coarse_Slicer(ode, 50)
    zs = rng.standard_normal((6, 120))
    r = api.Traj_sim_SIRS(init, np.tile(co, (6, 1, 1)), {"A": np.tile(kf["A"], (6, 1, 1, 1)), "Sigma": np.tile(kf["Sigma"], (6, 1, 1, 1))},
                          90.0, normals=zs)
    assert r["status"].shape == (6,) and len(set(r["status"].tolist())) == 1   # Sigma decides, not the normals
    print(f"Traj_sim_SIRS: chol(Sigma + 1e-13 I) exists for {ok_o}/{B} draws (oracle), {ok_d}/{B} (device), compared {both}")


def test_async_host_buffer_evaluation(changepoint):
    """lna_full_loglik_async: several calls in flight on the three rotating staging sets give, after lna_problem_sync, exactly
    what the synchronous call gives -- including when every staging set is reused (7 calls) and for a batch below the
    chunking threshold (falls back to the synchronous path)."""
    import ctypes as C
    from lnaphylodyn_b200 import _lib, api
    g = changepoint
    L = _lib.lib()
    pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    rng = np.random.default_rng(4)
    for B in (20000, 300):
        sets = []
        for s in range(7):
            par = np.tile(g.final["par"].ravel(), (B, 1))
            par[:, 2] *= np.exp(0.05 * rng.standard_normal(B))
            par[:, 3] *= np.exp(0.05 * rng.standard_normal(B))
            eps = np.ascontiguousarray(np.transpose(0.5 * rng.standard_normal((B, 40, 2)), (0, 2, 1)))
            sets.append((np.ascontiguousarray(par[:, 2:]), np.ascontiguousarray(par[:, :2]), eps, np.empty(B), np.empty(B, dtype=np.int32)))
        for prm, ini, eps, cl, st in sets:
            _lib.check(L.lna_full_loglik_async(pb.h, B, api._ptr(prm), api._ptr(ini), api._ptr(eps), api._ptr(cl), api._ptr(st)), "async")
        _lib.check(L.lna_problem_sync(pb.h), "sync")
        for prm, ini, eps, cl, st in sets:
            cl2, st2 = np.empty(B), np.empty(B, dtype=np.int32)
            _lib.check(L.lna_full_loglik(pb.h, B, api._ptr(prm), api._ptr(ini), api._ptr(eps), api._ptr(cl2), api._ptr(st2),
                                         None, None, None, None, None), "sync call")
            assert np.array_equal(cl, cl2) and np.array_equal(st, st2)
        assert len(np.unique(np.stack([s[3] for s in sets]), axis=0)) == 7
