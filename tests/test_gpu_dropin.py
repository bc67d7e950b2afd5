"""The reference-side binding, exercised end to end: integration/rcpp_shim.cpp -> C ABI -> sm_100a kernels.

`integration/rcpp_shim.cpp` is what a maintainer of the reference drops into the package's src/ (INTEGRATION.md): every
hot-path export with the reference's own C++ signature (src/LNA_functional.h:15-120), forwarding to the C ABI.  R and Rcpp
are not in the image, so it is compiled against the stand-in headers of oracle/shim/ -- the same ones the reference's own
sources compile against -- under the SAME harness (oracle/ref_harness.cpp):
    oracle/_ref/liblna_ref.so     harness + the reference's own function bodies          (CPU)
    oracle/_ref/liblna_dropin.so  harness + integration/rcpp_shim.cpp + the CUDA library  (GPU)
Both libraries are driven through the same Python wrappers with the same arguments and the same injected random streams.
Bar: values within 1e-10 relative (north_star); accept/reject decisions, slice-shrink counts, the number of uniforms and
normals the call leaves consumed in the session's generator, and reverted states identical.
"""
import numpy as np
import pytest

from conftest import Golden, relerr
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.dropin_available(), reason="oracle/_ref/liblna_dropin.so not built"),
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref/liblna_ref.so not built")]

TOL = 1e-10
PRIORS = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)]
ESS = [0, 1, 0, 0, 1, 1, 0]


@pytest.fixture(scope="module")
def dp():
    return ref.dropin()


def close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    assert float(np.max(np.abs(a - b))) <= tol * scale if a.size else True


def state(g, scale=0.0, seed=0):
    rng = np.random.default_rng(seed)
    par = g.final["par"].ravel().copy()
    eps = g.final["OriginTraj"].copy()
    if scale:
        par[2] *= np.exp(scale * rng.standard_normal())
        par[3] *= np.exp(scale * rng.standard_normal())
        par[4:-1] *= np.exp(0.2 * scale * rng.standard_normal(len(par) - 5))
        eps = eps + scale * rng.standard_normal(eps.shape)
    return par, eps


def chain_state(m, g, par, eps):
    npl = m.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
    lat = m.TransformTraj(npl["Ode"], eps, npl["FT"])
    return npl, lat, m.volz_loglik_nh2(g.init, lat, npl["betaN"], g.t_correct, g.x_i[2:4])


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_dropin_reproduces_cached_chain_state(dp, name):
    """The vignette's final MCMC state through the binding: what R + RcppArmadillo left in the knitr cache."""
    g = Golden(name)
    par, eps = state(g)
    npl, lat, cl = chain_state(dp, g, par, eps)
    assert relerr(npl["FT"]["A"], g.final["FT.A"]) < TOL
    assert np.max(np.abs(npl["FT"]["Sigma"] - g.final["FT.Sigma"])) < TOL * np.max(np.abs(g.final["FT.Sigma"]))
    assert relerr(npl["Ode"], g.final["Ode_Traj_coarse"]) < TOL
    assert relerr(npl["betaN"], g.final["betaN"].ravel()) < TOL
    assert relerr(lat, g.final["LatentTraj"]) < TOL
    assert cl == pytest.approx(float(g.final["coalLog"][0]), rel=TOL)
    for i in range(0, g.iters["par"].shape[0], 13):
        p_i, tr = g.iters["par"][i], g.iters["Trajectory"][i]
        bN = dp.betaTs(p_i[2:], tr[:, 0], g.x_r, g.x_i)
        assert dp.volz_loglik_nh2(g.init, tr, bN, g.t_correct, g.x_i[2:4]) == pytest.approx(g.iters["coalLog"][i], rel=TOL)


@pytest.mark.parametrize("model", ["SIR", "SEIR"])
def test_integrator_and_lna_moments(dp, changepoint, model):
    """a1-a7 through the binding: betaTs, param_transform, ODE_rk45(_stop), SigmaF, KF_param(_chol), Ode_Coarse_Slicer,
    New_Param_List."""
    g = changepoint
    rng = np.random.default_rng(3)
    nparam = 3 if model == "SEIR" else 2
    x_i = [39, nparam, 0, nparam - 1]
    param = np.concatenate([[2.3, 0.4, 0.21] if nparam == 3 else [2.3, 0.21], np.exp(0.05 * rng.standard_normal(39)), [20.0]])
    initial = np.array([1e6, 20.0, 10.0] if model == "SEIR" else [1e6, 10.0])
    t = g.times
    kw = dict(model=model)
    close(dp.betaTs(param, t[::50], g.x_r, x_i), ref.betaTs(param, t[::50], g.x_r, x_i), 1e-14)
    for tt in (0.0, 2.5, 33.3, 100.0):
        close(dp.param_transform(tt, param, g.x_r, x_i), ref.param_transform(tt, param, g.x_r, x_i), 1e-14)
    ode_d, ode_r = dp.ODE_rk45(initial, t, param, g.x_r, x_i, **kw), ref.ODE_rk45(initial, t, param, g.x_r, x_i, **kw)
    close(ode_d, ode_r)
    assert dp.ODE_rk45_stop(initial, t, param, g.x_r, x_i, tol=50.0, **kw) == ref.ODE_rk45_stop(initial, t, param, g.x_r, x_i,
                                                                                              tol=50.0, **kw)
    seg = ode_r[100:150]
    a, b = dp.SigmaF(seg, param, g.x_r, x_i, **kw), ref.SigmaF(seg, param, g.x_r, x_i, **kw)
    close(a["expF"], b["expF"])
    close(a["Sigma"], b["Sigma"])
    for fn in ("KF_param", "KF_param_chol"):
        a, b = getattr(dp, fn)(ode_r, param, 50, g.x_r, x_i, **kw), getattr(ref, fn)(ode_r, param, 50, g.x_r, x_i, **kw)
        close(a["A"], b["A"])
        close(a["Sigma"], b["Sigma"])
    assert np.array_equal(dp.Ode_Coarse_Slicer(ode_r, 50), ref.Ode_Coarse_Slicer(ode_r, 50))
    a = dp.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    b = ref.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    close(a["FT"]["A"], b["FT"]["A"])
    close(a["FT"]["Sigma"], b["FT"]["Sigma"])
    close(a["Ode"], b["Ode"])
    close(a["betaN"], b["betaN"], 1e-14)


def test_cholesky_failure_is_the_reference_exception(dp, changepoint):
    """KF_param_chol's std::invalid_argument (:663) comes out of the binding for the same input."""
    g = changepoint
    par, _ = state(g)
    bad = par.copy()
    bad[2] = 1e9
    for m in (ref, dp):
        with pytest.raises(ValueError):
            m.New_Param_List(bad[2:], bad[:2], g.gridsize, g.times, g.x_r, g.x_i)


def test_log_scale_through_the_binding(dp, changepoint):
    """transX = "log" (SIR): integrator, LNA moments and likelihood through the binding against the reference's own code;
    SEIR on the log scale -- never used by the package -- is refused with an R error (Rcpp::stop), not answered by a fallback."""
    g = changepoint
    rng = np.random.default_rng(3)
    x_i = [39, 2, 0, 1]
    param = np.concatenate([[2.3, 0.21], np.exp(0.05 * rng.standard_normal(39)), [20.0]])
    initial, t = np.log([1e6, 10.0]), g.times[:401]
    kw = dict(model="SIR", transX="log")
    ode_d, ode_r = dp.ODE_rk45(initial, t, param, g.x_r, x_i, **kw), ref.ODE_rk45(initial, t, param, g.x_r, x_i, **kw)
    close(ode_d, ode_r)
    a, b = dp.KF_param(ode_r, param, 50, g.x_r, x_i, **kw), ref.KF_param(ode_r, param, 50, g.x_r, x_i, **kw)
    close(a["A"], b["A"])
    close(a["Sigma"], b["Sigma"], 1e-9)
    a, b = dp.SigmaF(ode_r[100:150], param, g.x_r, x_i, **kw), ref.SigmaF(ode_r[100:150], param, g.x_r, x_i, **kw)
    close(a["expF"], b["expF"])
    close(a["Sigma"], b["Sigma"], 1e-9)
    lt = np.column_stack([g.final["LatentTraj"][:, 0], np.log(np.abs(g.final["LatentTraj"][:, 1:]) + 1.0)])
    bn = g.final["betaN"].ravel()
    assert dp.volz_loglik_nh2(g.init, lt, bn, g.t_correct, [0, 1], "log") == pytest.approx(
        ref.volz_loglik_nh2(g.init, lt, bn, g.t_correct, [0, 1], "log"), rel=TOL)
    with pytest.raises(RuntimeError, match="transX"):
        dp.ODE_rk45(np.log([1e6, 20.0, 10.0]), t, np.concatenate([[2.3, 0.4, 0.21], param[2:]]), g.x_r, [39, 3, 0, 2],
                    model="SEIR", transX="log")


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_trajectory_and_likelihood_functions(dp, name):
    """a8-a9: TransformTraj, Traj_sim_general_noncentral, volz_loglik_nh2/nh3 (+ sentinels), coal_loglik3, coal_loglik."""
    g = Golden(name)
    rng = np.random.default_rng(11)
    for s in range(4):
        par, eps = state(g, 0.05 * s, seed=s)
        npl = ref.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
        tr_d, tr_r = dp.TransformTraj(npl["Ode"], eps, npl["FT"]), ref.TransformTraj(npl["Ode"], eps, npl["FT"])
        close(tr_d, tr_r, 1e-13)
        z = rng.standard_normal(80)
        a = dp.Traj_sim_general_noncentral(npl["Ode"], npl["FT"], g.t_correct, z)
        b = ref.Traj_sim_general_noncentral(npl["Ode"], npl["FT"], g.t_correct, z)
        close(a["SimuTraj"], b["SimuTraj"], 1e-13)
        assert np.array_equal(a["OriginTraj"], b["OriginTraj"])
        assert a["logOrigin"] == pytest.approx(b["logOrigin"], rel=1e-13)
        assert a["logMultiNorm"] == pytest.approx(b["logMultiNorm"], rel=TOL)
        for traj in (tr_r, b["SimuTraj"]):
            for fn in ("volz_loglik_nh2", "volz_loglik_nh3"):
                va = getattr(dp, fn)(g.init, traj, npl["betaN"], g.t_correct, [0, 1])
                vb = getattr(ref, fn)(g.init, traj, npl["betaN"], g.t_correct, [0, 1])
                assert (np.isnan(va) and np.isnan(vb)) or va == pytest.approx(vb, rel=TOL)
            for lam in (1.0, 250.0):
                assert dp.coal_loglik3(g.init, traj, g.t_correct, lam, 1) == pytest.approx(
                    ref.coal_loglik3(g.init, traj, g.t_correct, lam, 1), rel=TOL, nan_ok=True)
        lt = tr_r.copy()
        lt[:, 1:] = np.log(np.abs(lt[:, 1:]) + 1.0)
        assert dp.coal_loglik(g.init, lt, g.t_correct, 3.0) == pytest.approx(ref.coal_loglik(g.init, lt, g.t_correct, 3.0), rel=TOL)
    neg = tr_r.copy()
    neg[5, 2] = -1.0
    assert dp.volz_loglik_nh2(g.init, neg, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL
    nan = tr_r.copy()
    nan[3, 1] = np.nan
    assert dp.volz_loglik_nh2(g.init, nan, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL


def test_update_param_decisions(dp, changepoint):
    """a10: same accept/reject decision and same accepted state for the same uniform."""
    g = changepoint
    rng = np.random.default_rng(5)
    coal_log = float(g.final["coalLog"][0])
    n_acc = 0
    for s in range(24):
        par, _ = state(g, 0.01, seed=100 + s)
        eps = g.final["OriginTraj"]
        u = rng.uniform()
        args = (par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, coal_log, 0.1, g.t_correct)
        a, b = dp.Update_Param(*args, unif=u), ref.Update_Param(*args, unif=u)
        assert bool(a["accept"]) == bool(b["accept"])
        if a["accept"]:
            n_acc += 1
            assert a["coalLog"] == pytest.approx(b["coalLog"], rel=TOL)
            close(a["LatentTraj"], b["LatentTraj"])
            close(a["FT"]["Sigma"], b["FT"]["Sigma"])
    assert 0 < n_acc < 24


def test_param_slice_updates(dp, changepoint):
    """a11: Param_Slice_update, Param_Slice_update_all2."""
    g = changepoint
    par, _ = state(g, 0.05, seed=2)
    rng = np.random.default_rng(8)
    for ess in (ESS, [1, 1, 1, 0, 1, 1, 0], [0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0, 0], [0, 0, 1, 0, 0, 0, 0]):
        th, nu = rng.uniform(-6, 6), rng.standard_normal(43)
        a = dp.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, PRIORS)
        b = ref.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, PRIORS)
        close(a, b, 1e-13)
        assert np.array_equal(a == par, b == par)
    th, nu = 0.7, rng.standard_normal(39)
    close(dp.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nu), ref.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nu), 1e-13)


def same_slice_result(a, b, keys):
    # what the call leaves consumed in the session's generator is what the unmodified package would leave
    assert a["n_uniforms"] == b["n_uniforms"] and a["n_normals"] == b["n_normals"]
    for k in keys:
        if np.ndim(a[k]) == 0:
            assert a[k] == pytest.approx(b[k], rel=TOL), k
        else:
            close(a[k], b[k])


def test_eslice_par_general_step_for_step(dp, changepoint):
    """a12: same number of shrinks, same accepted candidate, same state, same generator position afterwards."""
    g = changepoint
    shrinks = []
    for s in range(16):
        par, eps = state(g, 0.03 * (s % 4), seed=40 + s)
        cl = chain_state(ref, g, par, eps)[2]
        rng = np.random.default_rng([9, s])
        kw = dict(normals=rng.standard_normal(43), uniforms=rng.uniform(size=22))
        args = (par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, cl, g.t_correct)
        a, b = dp.ESlice_par_General(*args, **kw), ref.ESlice_par_General(*args, **kw)
        same_slice_result(a, b, ("par", "CoalLog", "LatentTraj", "betaN", "OdeTraj", "OriginTraj", "logOrigin"))
        close(a["FT"]["A"], b["FT"]["A"])
        close(a["FT"]["Sigma"], b["FT"]["Sigma"])
        shrinks.append(b["n_uniforms"] - 2)
    assert max(shrinks) >= 5 and min(shrinks) <= 3


def test_eslice_par_general_cap_and_failed_proposals(dp, changepoint):
    """The 20-shrink cap (:1609-1622) and proposals whose factorisation fails inside the loop (:1641-1651)."""
    g = changepoint
    par, eps = state(g)
    rng = np.random.default_rng(77)
    kw = dict(normals=rng.standard_normal(43), uniforms=rng.uniform(size=22))
    args = (par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, +1e6, g.t_correct)
    a, b = dp.ESlice_par_General(*args, **kw), ref.ESlice_par_General(*args, **kw)
    assert a["n_uniforms"] == b["n_uniforms"] == 21
    assert np.array_equal(a["par"], par) and np.array_equal(b["par"], par)
    assert a["CoalLog"] == pytest.approx(b["CoalLog"], rel=TOL)
    wild = [(1.0, 1.0), (0.7, 40.0), (-1.7, 0.1), (3.0, 0.2)]
    hit = 0
    cl = chain_state(ref, g, par, eps)[2]
    for s in range(8):
        rng = np.random.default_rng([78, s])
        kw = dict(normals=rng.standard_normal(43) * 3.0, uniforms=rng.uniform(size=22))
        args = (par, g.times, eps, wild, g.x_r, g.x_i, g.init, g.gridsize, ESS, cl, g.t_correct)
        a, b = dp.ESlice_par_General(*args, **kw), ref.ESlice_par_General(*args, **kw)
        same_slice_result(a, b, ("par", "CoalLog", "LatentTraj"))
        hit += int(b["n_uniforms"] > 4)
    assert hit > 0


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_eslice_general_nc_step_for_step(dp, name):
    """a13: the trajectory slice loop, with the caller's coal_log and with the -99999999 'recompute' default."""
    g = Golden(name)
    for s in range(12):
        par, eps = state(g, 0.02 * (s % 3), seed=60 + s)
        npl, _, cl0 = chain_state(ref, g, par, eps)
        rng = np.random.default_rng([13, s])
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        cl = cl0 if s % 2 == 0 else -99999999.0
        args = (eps, npl["Ode"], npl["FT"], par[:2], g.init, npl["betaN"], g.t_correct, 1.0, cl, g.gridsize, True, "SIR")
        a, b = dp.ESlice_general_NC(*args, **kw), ref.ESlice_general_NC(*args, **kw)
        same_slice_result(a, b, ("LatentTraj", "OriginTraj", "logOrigin", "CoalLog"))
    args = (eps, npl["Ode"], npl["FT"], par[:2], g.init, npl["betaN"], g.t_correct, 1.0, 1e6, g.gridsize, True, "SIR")
    a, b = dp.ESlice_general_NC(*args, **kw), ref.ESlice_general_NC(*args, **kw)
    assert a["n_uniforms"] == b["n_uniforms"] == 22
    assert np.array_equal(a["OriginTraj"], eps) and np.array_equal(b["OriginTraj"], eps)
    assert a["CoalLog"] == pytest.approx(b["CoalLog"], rel=TOL)


def test_eslice_change_points_step_for_step(dp, constant):
    """a14: the change-point-only slice loop of General_MCMC2 (uniforms u, theta first, then the normals)."""
    g = constant
    for s in range(10):
        par, eps = state(g, 0.02 * (s % 3), seed=80 + s)
        cl = chain_state(ref, g, par, eps)[2]
        rng = np.random.default_rng([14, s])
        kw = dict(normals=rng.standard_normal(39), uniforms=rng.uniform(size=22))
        args = (par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, cl, g.t_correct)
        a, b = dp.ESlice_change_points(*args, **kw), ref.ESlice_change_points(*args, **kw)
        same_slice_result(a, b, ("param", "CoalLog", "LatentTraj", "betaN", "OdeTraj", "LogChProb"))


def test_centred_parametrisation(dp, changepoint):
    """f3: log_like_traj_general2, Traj_sim_general3, Traj_sim_ezG2, ESlice_general2."""
    g = changepoint
    par, _ = state(g)
    param, initial = par[2:], par[:2]
    ode = ref.ODE_rk45(initial, g.times, param, g.x_r, g.x_i)
    kf = ref.KF_param(ode, param, 50, g.x_r, g.x_i)
    co = ref.Ode_Coarse_Slicer(ode, 50)
    betaN = ref.betaTs(param, co[:, 0], g.x_r, g.x_i)
    rng = np.random.default_rng(12)
    for s in range(6):
        z = rng.standard_normal(80)
        a, b = dp.Traj_sim_general3(co, kf, g.t_correct, normals=z), ref.Traj_sim_general3(co, kf, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], b["SimuTraj"][:, 1:])
        assert a["loglike"] == pytest.approx(b["loglike"], rel=1e-9)
        la = dp.log_like_traj_general2(b["SimuTraj"], co, kf, 50, g.t_correct)
        lb = ref.log_like_traj_general2(b["SimuTraj"], co, kf, 50, g.t_correct)
        assert la == pytest.approx(lb, rel=1e-9)
        a = dp.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
        c = ref.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], c["SimuTraj"][:, 1:])
        assert a["loglike"] == pytest.approx(c["loglike"], rel=1e-9)
        f_cur = b["SimuTraj"]
        if np.min(f_cur[:, 1:]) < 0:
            continue
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        args = (f_cur, co, kf, initial, g.init, betaN, g.t_correct, 10.0, 1, 50, True, "SIR")
        a, b2 = dp.ESlice_general2(*args, **kw), ref.ESlice_general2(*args, **kw)
        assert a["n_uniforms"] == b2["n_uniforms"]
        close(a["newTraj"], b2["newTraj"])


def test_seasonal_sirs_variant(dp):
    """row 13: ODE2, SIRS_KOM_Filter, log_like_trajSIRS of src/SIRS_period.cpp."""
    t = np.linspace(0, 100, 2001)
    param, period = [3e-7, 0.2, 0.05, 0.3], 40.0
    initial = [1e6, 10.0, 0.0]
    a, b = dp.ODE2(initial, t, param, period), ref.ODE2(initial, t, param, period)
    close(a, b)
    fa, fb = dp.SIRS_KOM_Filter(b, param, 50, period), ref.SIRS_KOM_Filter(b, param, 50, period)
    close(fa["A"], fb["A"])
    close(fa["Sigma"], fb["Sigma"])
    co = b[::50]
    rng = np.random.default_rng(6)
    X = co.copy()
    X[1:, 1:] += rng.standard_normal((40, 3)) * np.sqrt(np.abs(np.einsum("iik->ik", fb["Sigma"]).T)) * 0.5
    la, lb = dp.log_like_trajSIRS(X, co, fb, 50, 90.0), ref.log_like_trajSIRS(X, co, fb, 50, 90.0)
    assert la == pytest.approx(lb, rel=1e-6)  # pseudo-inverse of a rank-2 covariance: eigen-solver dependent digits


def _driver_over(mod):
    d = ref.driver()
    d.orc = mod
    return d


def test_r_level_iterations_over_the_binding(dp, changepoint):
    """f1: General_MCMC_with_ESlice's iteration body (the R code restated in oracle/driver.py) composed from the BINDING's
    exports -- i.e. the unmodified R driver running on the modified package -- against the same body over the reference's
    own functions: the two chains stay together over a fixed horizon."""
    g = changepoint
    rd, dd = _driver_over(ref), _driver_over(dp)
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    sa, sb = dd.init_state(g, par0, g.setting("control.traj")), rd.init_state(g, par0, g.setting("control.traj"))
    opts = rd.opts_from_golden(g)
    rng = np.random.default_rng(31)
    for it in range(10):
        nrm, unf = rng.standard_normal(43 + 80), rng.uniform(size=47)
        sa = dd.eslice_iteration(sa, g, opts, nrm, unf)
        sb = rd.eslice_iteration(sb, g, opts, nrm, unf)
        close(sa["par"], sb["par"], 1e-9)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-9)
        close(sa["LatentTraj"], sb["LatentTraj"], 1e-9)
        close(sa["OriginTraj"], sb["OriginTraj"], 1e-9)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]


def test_mcmc2_iterations_over_the_binding(dp, constant):
    """General_MCMC2's iteration body over the binding's exports."""
    g = constant
    rd, dd = _driver_over(ref), _driver_over(dp)
    par0, eps0 = state(g)
    sa, sb = dd.init_state(g, par0, eps0), rd.init_state(g, par0, eps0)
    opts = rd.mcmc2_opts_from_golden(g)
    rng = np.random.default_rng(32)
    for it in range(6):
        nrm, unf = rng.standard_normal(39 + 80), rng.uniform(size=54)
        sa = dd.mcmc2_iteration(sa, g, opts, nrm, unf)
        sb = rd.mcmc2_iteration(sb, g, opts, nrm, unf)
        close(sa["par"], sb["par"], 1e-9)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-9)
        close(sa["LatentTraj"], sb["LatentTraj"], 1e-9)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]
