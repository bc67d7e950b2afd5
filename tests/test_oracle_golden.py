"""Pin the CPU oracle against the reference's own shipped outputs (knitr caches -> tests/golden).

These are the golden vectors of SURVEY.md section 8c: the final chain state of both simulation vignettes
(ODE -> FT -> TransformTraj -> betaTs -> volz_loglik_nh2 end to end) and 2x104 per-iteration
(par, Trajectory) -> coalLog known answers.  Tolerance 1e-12 relative (observed <= 1e-15).
"""
import numpy as np
import pytest

from conftest import relerr
from oracle import oracle as orc

TOL = 1e-12


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_final_state_full_chain(name, request):
    g = request.getfixturevalue(name)
    par = g.final["par"].ravel()
    p = 2
    res = orc.New_Param_List(par[p:], par[:p], g.gridsize, g.times, g.x_r, g.x_i)
    assert relerr(res["Ode"], g.final["Ode_Traj_coarse"]) < TOL
    assert relerr(res["FT"]["A"], g.final["FT.A"]) < TOL
    assert np.max(np.abs(res["FT"]["Sigma"] - g.final["FT.Sigma"])) < 1e-11  # holds L (A.4); has exact zeros
    assert relerr(res["betaN"], g.final["betaN"].ravel()) < TOL
    lat = orc.TransformTraj(res["Ode"], g.final["OriginTraj"], res["FT"])
    assert relerr(lat[:, 1:], g.final["LatentTraj"][:, 1:]) < 1e-11
    ll = orc.volz_loglik_nh2(g.init, lat, res["betaN"], g.t_correct, g.x_i[2:4])
    assert abs(ll - float(g.final["coalLog"][0])) < TOL * abs(ll)
    # same through the stored LatentTraj
    ll2 = orc.volz_loglik_nh2(g.init, g.final["LatentTraj"], g.final["betaN"].ravel(), g.t_correct, g.x_i[2:4])
    assert abs(ll2 - float(g.final["coalLog"][0])) < TOL * abs(ll2)


def test_changepoint_anchor_values(changepoint):
    """The literal anchors quoted in SURVEY.md section 8c."""
    g = changepoint
    par = g.final["par"].ravel()
    assert par[:4] == pytest.approx([1e6, 1.0, 2.4510315651288463, 0.17156787370831236], rel=1e-15)
    res = orc.New_Param_List(par[2:], par[:2], 50, g.times, g.x_r, g.x_i)
    A0 = res["FT"]["A"][:, :, 0].ravel(order="F")
    L0 = res["FT"]["Sigma"][:, :, 0].ravel(order="F")
    assert A0 == pytest.approx([0.9999977375953129, 1.9304516030255832e-06, -1.4462694640797895,
                                1.856203003051246], rel=1e-13)
    assert L0 == pytest.approx([2.002356962910523, -1.812169895709815, 0, 0.6628898345345969], rel=1e-13)
    assert float(g.final["coalLog"][0]) == -2642.3217487475567


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_per_iteration_known_answers(name, request):
    """(par_i, Trajectory_i) -> coalLog_i for every stored iteration: pins betaTs + volz_loglik_nh2."""
    g = request.getfixturevalue(name)
    ini = orc.Init(g.init)
    worst = 0.0
    for par, traj, cl in zip(g.iters["par"], g.iters["Trajectory"], g.iters["coalLog"]):
        betaN = orc.betaTs(par[2:], traj[:, 0], g.x_r, g.x_i)
        ll = orc.volz_loglik_nh2(ini, traj, betaN, g.t_correct, g.x_i[2:4])
        worst = max(worst, abs(ll - cl) / abs(cl))
    assert worst < TOL, worst


def test_logorigin_quirk(changepoint):
    """ESlice_general_NC's logOrigin omits the last row (A.13): cache value -36.454165."""
    eps = changepoint.final["OriginTraj"]
    assert -0.5 * np.sum(eps[:-1] ** 2) == pytest.approx(float(changepoint.final["logOrigin"][0]), rel=1e-13)


def test_chprobs(changepoint):
    """General_MCMC_with_ESlice never refreshes chprobs: the cache holds the value from
    MCMC_initialize_general (general_change_point.R:552) at control$ch=0.975, hyper=20."""
    ch, hyper = changepoint.setting("control.ch"), float(changepoint.setting("control.hyper")[0])
    assert -0.5 * np.sum((np.log(ch) * hyper) ** 2) == pytest.approx(float(changepoint.final["chprobs"][0]), rel=1e-12)


def test_oracle_matches_reference_as_written_vectors(changepoint, constant):
    """tests/golden/ref_as_written.npz = outputs of the reference's own code (compiled verbatim on oracle/shim/,
    tests/golden/make_ref_golden.py).  The oracle must reproduce them; runs without /root/reference."""
    import os
    from conftest import GOLDEN
    from oracle import driver
    R = np.load(os.path.join(GOLDEN, "ref_as_written.npz"))
    for tag, g in (("full_cp", changepoint), ("full_co", constant)):
        par, eps = R[tag + ".par"], R[tag + ".eps"]
        cfg, ini = orc.Cfg(g.x_r, g.x_i, par.shape[1] - 2), orc.Init(g.init)
        for b in range(0, par.shape[0], 3):
            o = orc.full_loglik(cfg, ini, par[b, 2:], par[b, :2], g.times, g.gridsize, eps[b], g.t_correct)
            assert o["coalLog"] == pytest.approx(R[tag + ".coalLog"][b], rel=1e-12)
            assert relerr(o["LatentTraj"], R[tag + ".LatentTraj"][b]) < 1e-11
    g = changepoint
    par, eps = g.final["par"].ravel(), g.final["OriginTraj"]
    pri, ess = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)], [0, 1, 0, 0, 1, 1, 0]
    for b in range(R["esp.normals"].shape[0]):
        o = orc.ESlice_par_General(par, g.times, eps, pri, g.x_r, g.x_i, g.init, g.gridsize, ess, float(g.final["coalLog"][0]),
                                   g.t_correct, normals=R["esp.normals"][b], uniforms=R["esp.uniforms"][b])
        assert o["n_uniforms"] == R["esp.n_uniforms"][b]
        assert relerr(o["par"], R["esp.par"][b]) < 1e-12 and o["CoalLog"] == pytest.approx(R["esp.CoalLog"][b], rel=1e-12)
    # the chain trace
    nrm, unf = R["chain.normals"], R["chain.uniforms"]
    st = [driver.init_state(g, R["chain.par0"], g.setting("control.traj")) for _ in range(nrm.shape[1])]
    for it in range(nrm.shape[0]):
        for c in range(nrm.shape[1]):
            st[c] = driver.eslice_iteration(st[c], g, driver.opts_from_golden(g, update_traj=it >= 2), nrm[it, c], unf[it, c])
            assert relerr(st[c]["par"], R["chain.par"][it, c]) < 1e-11
            assert st[c]["coalLog"] == pytest.approx(R["chain.coalLog"][it, c], rel=1e-11)
