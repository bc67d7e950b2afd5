"""The C oracle against THE REFERENCE CODE AS WRITTEN (CPU only).

`oracle/_ref/liblna_ref.so` is the reference's own sources compiled verbatim from /root/reference/src against the
stand-in Armadillo/Rcpp headers of oracle/shim/ (oracle/Makefile target `ref`; built by __graft_entry__.build()
in the container that has /root/reference, shipped prebuilt to the GPU box).  These tests pin
  * the stand-in headers: the reference code compiled on them reproduces the numbers the REAL R + RcppArmadillo
    build left in the package's knitr caches, to the last bit for coalLog;
  * the oracle: every function of SURVEY.md section 8(a) -- including the Metropolis step and the three
    elliptical-slice loops, whose accept/shrink sequences no reference artefact covers -- agrees with the
    reference code on the same inputs and the same pre-drawn random streams.
Tolerance: 1e-12 relative on values (the two implementations differ by operation order only); decisions,
consumed-stream counts and reverted states must be identical.
"""
import numpy as np
import pytest

from conftest import Golden, relerr
from oracle import driver as odriver
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/liblna_ref.so not built and /root/reference absent")

TOL = 1e-12
PRIORS = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)]
ESS = [0, 1, 0, 0, 1, 1, 0]


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


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_reference_on_shim_reproduces_cached_chain_state(name):
    """ODE_rk45 -> KF_param_chol -> Ode_Coarse_Slicer -> betaTs -> TransformTraj -> volz_loglik_nh2 of the cached final
    MCMC state, run by the reference's own code: equal to what R + RcppArmadillo computed."""
    g = Golden(name)
    par, eps = state(g)
    r = ref.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
    assert relerr(r["FT"]["A"], g.final["FT.A"]) < 1e-13
    assert np.max(np.abs(r["FT"]["Sigma"] - g.final["FT.Sigma"])) < 1e-12 * np.max(np.abs(g.final["FT.Sigma"]))
    assert relerr(r["Ode"], g.final["Ode_Traj_coarse"]) < 1e-13
    assert relerr(r["betaN"], g.final["betaN"].ravel()) < 1e-14
    traj = ref.TransformTraj(r["Ode"], eps, r["FT"])
    assert relerr(traj, g.final["LatentTraj"]) < 1e-12
    cl = ref.volz_loglik_nh2(g.init, traj, r["betaN"], g.t_correct, g.x_i[2:4])
    assert cl == float(g.final["coalLog"][0])  # bit-exact, also the summation order
    # per-iteration known answers (par_i, Trajectory_i) -> l1_i
    for i in range(0, g.iters["par"].shape[0], 7):
        p_i = g.iters["par"][i]
        tr = g.iters["Trajectory"][i]
        bN = ref.betaTs(p_i[2:], tr[:, 0], g.x_r, g.x_i)
        assert ref.volz_loglik_nh2(g.init, tr, bN, g.t_correct, g.x_i[2:4]) == pytest.approx(g.iters["coalLog"][i], rel=1e-13)


@pytest.mark.parametrize("model,transX", [("SIR", "standard"), ("SIR", "log"), ("SEIR", "standard"), ("SEIR", "log"),
                                          ("SEIR2", "standard")])
def test_integrator_and_lna_moments(changepoint, model, transX):
    """a1-a7: betaTs, param_transform, ODE_rk45(_stop), SigmaF, KF_param, KF_param_chol, Ode_Coarse_Slicer, New_Param_List."""
    g = changepoint
    rng = np.random.default_rng(3)
    nparam = 3 if model.startswith("SEIR") else 2
    x_i = [39, nparam, 0, nparam - 1]
    param = np.concatenate([[2.3, 0.4, 0.21][:nparam] if nparam == 3 else [2.3, 0.21],
                            np.exp(0.05 * rng.standard_normal(39)), [20.0]])
    p = orc.MODEL_P[model]
    if transX == "log":
        initial = np.log([1e6, 20.0, 10.0][:p] if p == 3 else [1e6, 10.0])
        t = g.times[:401]
    else:
        initial = np.array([1e6, 20.0, 10.0][:p] if model == "SEIR" else ([20.0, 10.0] if model == "SEIR2" else [1e6, 10.0]))
        t = g.times
    kw = dict(model=model, transX=transX)
    close(ref.betaTs(param, t[::50], g.x_r, x_i), orc.betaTs(param, t[::50], g.x_r, x_i), 1e-15)
    for tt in (0.0, 2.5, 33.3, 100.0):
        close(ref.param_transform(tt, param, g.x_r, x_i), orc.param_transform(tt, param, g.x_r, x_i), 1e-15)
    ode_r, ode_o = ref.ODE_rk45(initial, t, param, g.x_r, x_i, **kw), orc.ODE_rk45(initial, t, param, g.x_r, x_i, **kw)
    close(ode_r, ode_o, 1e-13)
    if model != "SEIR2":  # X1(x_i(3)) indexes a 2-vector with 2 there: an Armadillo bounds error in the reference
        tol = 50.0 if transX == "standard" else np.log(50.0)
        assert ref.ODE_rk45_stop(initial, t, param, g.x_r, x_i, tol=tol, **kw) == orc.ODE_rk45_stop(initial, t, param, g.x_r, x_i,
                                                                                                tol=tol, **kw)
    if model == "SEIR2" or (model == "SEIR" and transX == "log"):
        return  # SigmaF: SEIR2 multiplies non-conformant matrices (3x2 stoichiometry); SEIR/log is never used by the package
    seg = ode_o[100:150]
    a, b = ref.SigmaF(seg, param, g.x_r, x_i, **kw), orc.SigmaF(seg, param, g.x_r, x_i, **kw)
    close(a["expF"], b["expF"])
    close(a["Sigma"], b["Sigma"])
    a, b = ref.KF_param(ode_o, param, 50, g.x_r, x_i, **kw), orc.KF_param(ode_o, param, 50, g.x_r, x_i, **kw)
    close(a["A"], b["A"])
    close(a["Sigma"], b["Sigma"])
    if transX == "log":
        return
    a, b = ref.KF_param_chol(ode_o, param, 50, g.x_r, x_i, **kw), orc.KF_param_chol(ode_o, param, 50, g.x_r, x_i, **kw)
    close(a["A"], b["A"])
    close(a["Sigma"], b["Sigma"], 1e-11)
    assert np.array_equal(ref.Ode_Coarse_Slicer(ode_o, 50), orc.Ode_Coarse_Slicer(ode_o, 50))
    a = ref.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    b = orc.New_Param_List(param, initial, 50, t, g.x_r, x_i, **kw)
    close(a["FT"]["A"], b["FT"]["A"])
    close(a["FT"]["Sigma"], b["FT"]["Sigma"], 1e-11)
    close(a["Ode"], b["Ode"], 1e-13)
    close(a["betaN"], b["betaN"], 1e-15)


def test_cholesky_failure_is_the_same_event(changepoint):
    """KF_param_chol throws (reference) exactly where the oracle reports ORC_CHOL_FAIL."""
    g = changepoint
    par, _ = state(g)
    bad = par.copy()
    bad[2] = 1e9  # absurd R0: the epidemic burns out in one step, Sigma goes non-positive / NaN
    for m in (ref, orc):
        with pytest.raises(ValueError):
            m.New_Param_List(bad[2:], bad[:2], g.gridsize, g.times, g.x_r, g.x_i)


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_trajectory_and_likelihood_functions(name):
    """a8-a9: TransformTraj, Traj_sim_general_noncentral, volz_loglik_nh2/nh3 (+ sentinels), coal_loglik3, coal_loglik."""
    g = Golden(name)
    rng = np.random.default_rng(11)
    for s in range(6):
        par, eps = state(g, 0.05 * s, seed=s)
        npl = orc.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
        tr_r, tr_o = ref.TransformTraj(npl["Ode"], eps, npl["FT"]), orc.TransformTraj(npl["Ode"], eps, npl["FT"])
        close(tr_r, tr_o, 1e-14)
        z = rng.standard_normal(80)
        a = ref.Traj_sim_general_noncentral(npl["Ode"], npl["FT"], g.t_correct, z)
        b = orc.Traj_sim_general_noncentral(npl["Ode"], npl["FT"], g.t_correct, z)
        close(a["SimuTraj"], b["SimuTraj"], 1e-14)
        assert np.array_equal(a["OriginTraj"], b["OriginTraj"])
        assert a["logOrigin"] == pytest.approx(b["logOrigin"], rel=1e-14)
        assert a["logMultiNorm"] == pytest.approx(b["logMultiNorm"], rel=1e-12)
        for traj in (tr_o, b["SimuTraj"]):
            for fn in ("volz_loglik_nh2", "volz_loglik_nh3"):
                va = getattr(ref, fn)(g.init, traj, npl["betaN"], g.t_correct, [0, 1])
                vb = getattr(orc, fn)(g.init, traj, npl["betaN"], g.t_correct, [0, 1])
                assert (np.isnan(va) and np.isnan(vb)) or va == pytest.approx(vb, rel=1e-13)
            for lam in (1.0, 250.0):
                assert ref.coal_loglik3(g.init, traj, g.t_correct, lam, 1) == pytest.approx(
                    orc.coal_loglik3(g.init, traj, g.t_correct, lam, 1), rel=1e-13, nan_ok=True)
        lt = tr_o.copy()
        lt[:, 1:] = np.log(np.abs(lt[:, 1:]) + 1.0)
        assert ref.coal_loglik(g.init, lt, g.t_correct, 3.0) == pytest.approx(orc.coal_loglik(g.init, lt, g.t_correct, 3.0),
                                                                               rel=1e-13)
        assert ref.volz_loglik_nh2(g.init, lt, npl["betaN"], g.t_correct, [0, 1], "log") == pytest.approx(
            orc.volz_loglik_nh2(g.init, lt, npl["betaN"], g.t_correct, [0, 1], "log"), rel=1e-13)
    # the two in-band sentinels
    neg = tr_o.copy()
    neg[5, 2] = -1.0
    assert ref.volz_loglik_nh2(g.init, neg, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL
    assert orc.volz_loglik_nh2(g.init, neg, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL
    nan = tr_o.copy()
    nan[3, 1] = np.nan
    assert ref.volz_loglik_nh2(g.init, nan, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL
    assert orc.volz_loglik_nh2(g.init, nan, npl["betaN"], g.t_correct, [0, 1]) == orc.SENTINEL


def test_update_param_decisions(changepoint):
    """a10: same accept/reject decision and same accepted state for the same uniform."""
    g = changepoint
    rng = np.random.default_rng(5)
    coal_log = float(g.final["coalLog"][0])
    n_acc = 0
    for s in range(24):
        par, eps = state(g, 0.01, seed=100 + s)
        eps = g.final["OriginTraj"]
        u = rng.uniform()
        a = ref.Update_Param(par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, coal_log, 0.1, g.t_correct, unif=u)
        b = orc.Update_Param(par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, coal_log, 0.1, g.t_correct, unif=u)
        assert bool(a["accept"]) == bool(b["accept"])
        if a["accept"]:
            n_acc += 1
            assert a["coalLog"] == pytest.approx(b["coalLog"], rel=1e-13)
            close(a["LatentTraj"], b["LatentTraj"])
            close(a["FT"]["Sigma"], b["FT"]["Sigma"], 1e-11)
    assert 0 < n_acc < 24


def test_param_slice_updates(changepoint):
    """a11: Param_Slice_update, Param_Slice_update_all2 for every ESS_vec pattern the drivers use."""
    g = changepoint
    par, _ = state(g, 0.05, seed=2)
    rng = np.random.default_rng(8)
    for ess in (ESS, [1, 1, 1, 0, 1, 1, 0], [0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0, 0], [0, 0, 1, 0, 0, 0, 0]):
        th, nu = rng.uniform(-6, 6), rng.standard_normal(43)
        a = ref.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, PRIORS)
        b = orc.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, PRIORS)
        close(a, b, 1e-14)
        assert np.array_equal(a == par, b == par)  # the untouched entries are the same entries
    th, nu = 0.7, rng.standard_normal(39)
    close(ref.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nu), orc.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nu), 1e-14)


def same_slice_result(a, b, keys):
    assert a["n_uniforms"] == b["n_uniforms"] and a["n_normals"] == b["n_normals"]
    for k in keys:
        if np.ndim(a[k]) == 0:
            assert a[k] == pytest.approx(b[k], rel=1e-12), k
        else:
            close(a[k], b[k], 1e-11 if k == "FT" else TOL)


def test_eslice_par_general_step_for_step(changepoint):
    """a12: the parameter slice loop -- same number of shrinks, same accepted candidate, same state."""
    g = changepoint
    shrinks = []
    for s in range(16):
        par, eps = state(g, 0.03 * (s % 4), seed=40 + s)
        st = odriver.init_state(g, par, eps)
        rng = np.random.default_rng([9, s])
        nrm, unf = rng.standard_normal(43), rng.uniform(size=22)
        kw = dict(normals=nrm, uniforms=unf)
        a = ref.ESlice_par_General(par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, st["coalLog"], g.t_correct, **kw)
        b = orc.ESlice_par_General(par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, st["coalLog"], g.t_correct, **kw)
        same_slice_result(a, b, ("par", "CoalLog", "LatentTraj", "betaN", "OdeTraj", "OriginTraj", "logOrigin"))
        close(a["FT"]["A"], b["FT"]["A"])
        close(a["FT"]["Sigma"], b["FT"]["Sigma"], 1e-11)
        shrinks.append(b["n_uniforms"] - 2)
    assert max(shrinks) >= 5 and min(shrinks) <= 3  # short and long shrink sequences were both exercised


def test_eslice_par_general_cap_and_failed_proposals(changepoint):
    """The 20-shrink cap (revert to the current parameters, :1609-1622) and proposals whose factorisation fails inside
    the loop (try/catch -> -1e7 -> keep shrinking, :1641-1651)."""
    g = changepoint
    par, eps = state(g)
    rng = np.random.default_rng(77)
    nrm, unf = rng.standard_normal(43), rng.uniform(size=22)
    kw = dict(normals=nrm, uniforms=unf)
    a = ref.ESlice_par_General(par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, +1e6, g.t_correct, **kw)
    b = orc.ESlice_par_General(par, g.times, eps, PRIORS, g.x_r, g.x_i, g.init, g.gridsize, ESS, +1e6, g.t_correct, **kw)
    assert a["n_uniforms"] == b["n_uniforms"] == 21 and b["status"] & 8
    assert np.array_equal(a["par"], par) and np.array_equal(b["par"], par)
    assert a["CoalLog"] == pytest.approx(b["CoalLog"], rel=1e-13)
    # huge prior scale on R0: the first wide-angle proposals blow the ODE up (chol failure), later ones are fine
    wild = [(1.0, 1.0), (0.7, 40.0), (-1.7, 0.1), (3.0, 0.2)]
    hit = 0
    for s in range(8):
        rng = np.random.default_rng([78, s])
        kw = dict(normals=rng.standard_normal(43) * 3.0, uniforms=rng.uniform(size=22))
        cl = odriver.init_state(g, par, eps)["coalLog"]
        a = ref.ESlice_par_General(par, g.times, eps, wild, g.x_r, g.x_i, g.init, g.gridsize, ESS, cl, g.t_correct, **kw)
        b = orc.ESlice_par_General(par, g.times, eps, wild, g.x_r, g.x_i, g.init, g.gridsize, ESS, cl, g.t_correct, **kw)
        same_slice_result(a, b, ("par", "CoalLog", "LatentTraj"))
        hit += int(b["n_uniforms"] > 4)
    assert hit > 0


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_eslice_general_nc_step_for_step(name):
    """a13: the trajectory slice loop, with the caller's coal_log and with the -99999999 'recompute' default."""
    g = Golden(name)
    for s in range(12):
        par, eps = state(g, 0.02 * (s % 3), seed=60 + s)
        st = odriver.init_state(g, par, eps)
        rng = np.random.default_rng([13, s])
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        cl = st["coalLog"] if s % 2 == 0 else -99999999.0
        args = (eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, 1.0, cl, g.gridsize, True, "SIR")
        a, b = ref.ESlice_general_NC(*args, **kw), orc.ESlice_general_NC(*args, **kw)
        same_slice_result(a, b, ("LatentTraj", "OriginTraj", "logOrigin", "CoalLog"))
    # cap: nothing can beat +1e6 -> 20 shrinks, f_prime = f_cur
    args = (eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, 1.0, 1e6, g.gridsize, True, "SIR")
    a, b = ref.ESlice_general_NC(*args, **kw), orc.ESlice_general_NC(*args, **kw)
    assert a["n_uniforms"] == b["n_uniforms"] == 22
    assert np.array_equal(a["OriginTraj"], eps) and np.array_equal(b["OriginTraj"], eps)
    assert a["CoalLog"] == pytest.approx(b["CoalLog"], rel=1e-13)


def test_eslice_general_nc_kingman_branch(changepoint):
    """volz = FALSE goes through LogTraj (SIR_LNA.cpp:40) and coal_loglik3, which logs the trajectory a second time:
    log(log I), NaN as soon as I < 1 -- and a NaN ends the loop (`NaN <= logy` is false) and is returned.  Reached only with
    likelihood != "volz" (no BASELINE config); oracle vs the reference code as written, step for step, over given
    coal_log values that end at once, shrink several times, hit the 20-cap (which recomputes with volz_loglik_nh2), and a
    trajectory pushed below one infective (NaN)."""
    g = changepoint
    seen_nan = seen_shrink = seen_cap = 0
    for s in range(14):
        par, eps = state(g, 0.02 * (s % 3), seed=160 + s)
        par = par.copy()
        par[1] = 0.5 if s >= 10 else 80.0   # < 1 infective at the start: log(log I) is NaN there; 80: finite everywhere
        st = odriver.init_state(g, par, eps)
        rng = np.random.default_rng([23, s])
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        lam = (1.0, 300.0)[s % 2]
        lat0 = np.array(st["LatentTraj"], order="F")
        base = orc.coal_loglik3(g.init, np.column_stack([lat0[:, 0], np.log(lat0[:, 1:])]), g.t_correct, lam, 1)
        cl = (base - 3.0, base + 0.5, 1e9, base)[s % 4] if np.isfinite(base) else -2000.0
        args = (eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, lam, cl, g.gridsize, False, "SIR")
        a, b = ref.ESlice_general_NC(*args, **kw), orc.ESlice_general_NC(*args, **kw)
        assert a["n_uniforms"] == b["n_uniforms"] and a["n_normals"] == b["n_normals"] == 80, s
        assert np.allclose(a["OriginTraj"], b["OriginTraj"], rtol=1e-13, atol=0) and np.allclose(a["LatentTraj"], b["LatentTraj"], rtol=1e-12)
        if np.isnan(a["CoalLog"]):
            assert np.isnan(b["CoalLog"])
            seen_nan += 1
        else:
            assert a["CoalLog"] == pytest.approx(b["CoalLog"], rel=1e-12)
        assert a["logOrigin"] == pytest.approx(b["logOrigin"], rel=1e-13)
        seen_shrink += a["n_uniforms"] > 3
        seen_cap += a["n_uniforms"] == 22
    assert seen_nan >= 1 and seen_shrink >= 2 and seen_cap >= 1, (seen_nan, seen_shrink, seen_cap)
    with pytest.raises(NotImplementedError):   # the -99999999 default in this branch treats the increments as a trajectory
        orc.ESlice_general_NC(eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, 300.0, -99999999.0,
                              g.gridsize, False, "SIR", normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))


def test_eslice_change_points_step_for_step(constant):
    """a14: the change-point-only slice loop of General_MCMC2 (uniforms u, theta first, then the normals)."""
    g = constant
    for s in range(10):
        par, eps = state(g, 0.02 * (s % 3), seed=80 + s)
        st = odriver.init_state(g, par, eps)
        rng = np.random.default_rng([14, s])
        kw = dict(normals=rng.standard_normal(39), uniforms=rng.uniform(size=22))
        args = (par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, st["coalLog"], g.t_correct)
        a, b = ref.ESlice_change_points(*args, **kw), orc.ESlice_change_points(*args, **kw)
        same_slice_result(a, b, ("param", "CoalLog", "LatentTraj", "betaN", "OdeTraj", "LogChProb"))


def test_centred_parametrisation(changepoint):
    """f3: log_like_traj_general2, Traj_sim_general3, Traj_sim_ezG2, ESlice_general2."""
    g = changepoint
    par, _ = state(g)
    param, initial = par[2:], par[:2]
    ode = orc.ODE_rk45(initial, g.times, param, g.x_r, g.x_i)
    kf = orc.KF_param(ode, param, 50, g.x_r, g.x_i)
    co = orc.Ode_Coarse_Slicer(ode, 50)
    betaN = orc.betaTs(param, co[:, 0], g.x_r, g.x_i)
    rng = np.random.default_rng(12)
    for s in range(6):
        z = rng.standard_normal(80)
        a, b = ref.Traj_sim_general3(co, kf, g.t_correct, normals=z), orc.Traj_sim_general3(co, kf, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], b["SimuTraj"][:, 1:], 1e-12)
        assert a["loglike"] == pytest.approx(b["loglike"], rel=1e-9)
        la = ref.log_like_traj_general2(b["SimuTraj"], co, kf, 50, g.t_correct)
        lb = orc.log_like_traj_general2(b["SimuTraj"], co, kf, 50, g.t_correct)
        assert la == pytest.approx(lb, rel=1e-9)
        a = ref.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
        c = orc.Traj_sim_ezG2(initial, g.times, param, 50, g.x_r, g.x_i, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], c["SimuTraj"][:, 1:], 1e-12)
        assert a["loglike"] == pytest.approx(c["loglike"], rel=1e-9)
        f_cur = b["SimuTraj"]
        if np.min(f_cur[:, 1:]) < 0:
            continue
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        args = (f_cur, co, kf, initial, g.init, betaN, g.t_correct, 10.0, 1, 50, True, "SIR")
        a, b2 = ref.ESlice_general2(*args, **kw), orc.ESlice_general2(*args, **kw)
        assert a["n_uniforms"] == b2["n_uniforms"]
        close(a["newTraj"], b2["newTraj"], 1e-12)


def test_seasonal_sirs_variant():
    """row 13: ODE2, SIRS_KOM_Filter, log_like_trajSIRS of src/SIRS_period.cpp."""
    t = np.linspace(0, 100, 2001)
    param, period = [3e-7, 0.2, 0.05, 0.3], 40.0
    initial = [1e6, 10.0, 0.0]
    a, b = ref.ODE2(initial, t, param, period), orc.ODE2(initial, t, param, period)
    close(a, b, 1e-13)
    fa, fb = ref.SIRS_KOM_Filter(b, param, 50, period), orc.SIRS_KOM_Filter(b, param, 50, period)
    close(fa["A"], fb["A"])
    close(fa["Sigma"], fb["Sigma"])
    co = b[::50]
    rng = np.random.default_rng(6)
    X = co.copy()
    X[1:, 1:] += rng.standard_normal((40, 3)) * np.sqrt(np.abs(np.einsum("iik->ik", fb["Sigma"]).T)) * 0.5
    la, lb = ref.log_like_trajSIRS(X, co, fb, 50, 90.0), orc.log_like_trajSIRS(X, co, fb, 50, 90.0)
    assert la == pytest.approx(lb, rel=1e-6)  # pseudo-inverse of a rank-2 covariance: eigen-solver dependent digits


def test_driver_iterations_step_for_step(changepoint):
    """f1: General_MCMC_with_ESlice's iteration body (MH gamma, no-op hyper MH, parameter ESS, trajectory ESS) composed
    from the reference's functions and from the oracle's: the two chains stay together over a fixed horizon."""
    g = changepoint
    rdriver = ref.driver()
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    sa, sb = rdriver.init_state(g, par0, g.setting("control.traj")), odriver.init_state(g, par0, g.setting("control.traj"))
    opts = odriver.opts_from_golden(g)
    rng = np.random.default_rng(31)
    for it in range(12):
        nrm, unf = rng.standard_normal(43 + 80), rng.uniform(size=47)
        sa = rdriver.eslice_iteration(sa, g, opts, nrm, unf)
        sb = odriver.eslice_iteration(sb, g, opts, nrm, unf)
        close(sa["par"], sb["par"], 1e-11)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-11)
        close(sa["LatentTraj"], sb["LatentTraj"], 1e-10)
        close(sa["OriginTraj"], sb["OriginTraj"], 1e-11)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]


def test_mcmc2_iterations_step_for_step(constant):
    """General_MCMC2's iteration body (three scalar MH entries, hyper MH, change-point ESS, trajectory ESS)."""
    g = constant
    rdriver = ref.driver()
    par0, eps0 = state(g)
    sa, sb = rdriver.init_state(g, par0, eps0), odriver.init_state(g, par0, eps0)
    opts = odriver.mcmc2_opts_from_golden(g)
    rng = np.random.default_rng(32)
    for it in range(8):
        nrm, unf = rng.standard_normal(39 + 80), rng.uniform(size=54)
        sa = rdriver.mcmc2_iteration(sa, g, opts, nrm, unf)
        sb = odriver.mcmc2_iteration(sb, g, opts, nrm, unf)
        close(sa["par"], sb["par"], 1e-11)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-11)
        close(sa["LatentTraj"], sb["LatentTraj"], 1e-10)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]


def test_general_entry_list_and_joint_stage_iterations(changepoint, constant):
    """The restated R iteration bodies added for the Ebola vignette's entry list (Update_Param_each_Norm with arbitrary
    entries, hyper = T) and for General_MCMC2's third stage (update_Param_joint, admcmc), composed from the reference's
    functions and from the oracle's: the two chains stay together."""
    rdriver = ref.driver()
    # --- General_MCMC_with_ESlice with entries (S0 log walk, R0 log walk, hyper log-normal walk)
    g = changepoint
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    sa, sb = rdriver.init_state(g, par0, g.setting("control.traj")), odriver.init_state(g, par0, g.setting("control.traj"))
    opts = dict(odriver.opts_from_golden(g), hyper=True,
                entries=[(0, 1, (13.8, 0.5), 0.2), (2, 1, (0.7, 0.5), 0.05), (-1, 0, (3.0, 0.2), 0.25)])
    rng = np.random.default_rng(41)
    for it in range(6):
        nrm, unf = rng.standard_normal(43 + 80), rng.uniform(size=52)
        sa = rdriver.eslice_iteration_general(sa, g, opts, nrm, unf)
        sb = odriver.eslice_iteration_general(sb, g, opts, nrm, unf)
        close(sa["par"], sb["par"], 1e-11)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-11)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]
    assert sb["par"][43] != par0[43] and sb["n_mh_accept"] > 0
    # --- General_MCMC2, stage i >= warmup2
    g = constant
    par0, eps0 = state(g)
    sa, sb = rdriver.init_state(g, par0, eps0), odriver.init_state(g, par0, eps0)
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    jopts = dict(odriver.mcmc2_opts_from_golden(g),
                 joint=[(1, 1, 0, tuple(prior[0][:2])), (2, 0, 1, tuple(prior[1][:2])), (3, 1, 0, tuple(prior[2][:2])),
                        (4, 0, 2, tuple(prior[4][:2]))])
    A = rng.standard_normal((4, 4)) * np.array([0.02, 0.004, 0.01, 0.3])[:, None]
    ev, vec = np.linalg.eigh(A @ A.T)
    T = vec * np.sqrt(np.maximum(ev, 0.0))[None, :]
    for it in range(6):
        nrm, unf = rng.standard_normal(39 + 80 + 4), rng.uniform(size=54)
        sa = rdriver.mcmc2_joint_iteration(sa, g, jopts, T, nrm, unf)
        sb = odriver.mcmc2_joint_iteration(sb, g, jopts, T, nrm, unf)
        close(sa["par"], sb["par"], 1e-11)
        assert sa["coalLog"] == pytest.approx(sb["coalLog"], rel=1e-11)
        assert sa["n_mh_accept"] == sb["n_mh_accept"]
    assert sb["n_mh_accept"] > 0
