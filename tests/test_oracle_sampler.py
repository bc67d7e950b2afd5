"""CPU-only checks of the oracle's samplers and of the reference's own testthat properties
(tests/testthat/testStructureCoal.R) restated on the oracle."""
import numpy as np
import pytest

from oracle import driver
from oracle import oracle as orc


def test_testthat_chol_2by2():
    """'Test Chol decomposition 2by2' (testStructureCoal.R:31-38)."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((2, 2))
    M = X @ X.T + np.eye(2)
    L = orc.chols(M)
    assert np.sum(np.abs(orc.inv2(M) - np.linalg.inv(M))) < 0.01
    assert L[0, 1] == 0
    assert np.sum(np.abs(L @ L.T - M)) < 1e-4


def test_testthat_sir_model(changepoint):
    """'Test SIR model' (testStructureCoal.R:41-51): (a) trailing params are ignored by ODE_rk45,
    (b) KF_param$A == KF_param_chol$A and Sigma = L L', (c) the non-centred / centred identity."""
    g = changepoint
    t = np.linspace(0, 100, 2001)
    x_r, x_i = [1e6, 20, 60], [2, 2, 0, 1]
    a = orc.ODE_rk45([1e6, 10], t, [2.2, 0.2, 0.8, 0.6], x_r, x_i)
    b = orc.ODE_rk45([1e6, 10], t, [2.2, 0.2, 0.8, 0.6, 7.0, 9.0], x_r, x_i)
    assert np.sum(np.abs(a - b)) == 0
    kf1 = orc.KF_param(a, [2.2, 0.2, 0.8, 0.6], 50, x_r, x_i)
    kf2 = orc.KF_param_chol(a, [2.2, 0.2, 0.8, 0.6], 50, x_r, x_i)
    assert np.array_equal(kf1["A"], kf2["A"])
    L1 = kf2["Sigma"][:, :, 0]
    assert np.sum(np.abs(kf1["Sigma"][:, :, 0] - L1 @ L1.T)) < 1e-6
    coarse = orc.Ode_Coarse_Slicer(a, 50)
    rng = np.random.default_rng(2)
    sim = orc.Traj_sim_general_noncentral(coarse, kf2, 150.0, rng.standard_normal(80))
    lhs = sim["logMultiNorm"] + sim["logOrigin"]
    rhs = orc.log_like_traj_general2(sim["SimuTraj"], coarse, kf1, 50, 150.0)
    assert lhs == pytest.approx(rhs, rel=1.5e-8)  # testthat's default tolerance; differs by the 1e-8 jitter


def test_eslice_par_general_consumption_and_invariants(changepoint):
    g = changepoint
    st = driver.init_state(g, g.final["par"].ravel(), g.final["OriginTraj"])
    assert st["coalLog"] == pytest.approx(float(g.final["coalLog"][0]), rel=1e-12)
    opts = driver.opts_from_golden(g)
    rng = np.random.default_rng(7)
    r = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init, 50,
                               opts["ESS_vec"], st["coalLog"], g.t_correct, normals=rng.standard_normal(43),
                               uniforms=rng.random(22))
    assert r["n_normals"] == 43 and r["n_uniforms"] == 1 + r["n_evals"] or r["status"] & 8
    assert r["par"][0] == st["par"][0] and r["par"][1] == st["par"][1]      # I0 frozen (ESS_vec[1] = 0)
    assert r["par"][3] == st["par"][3]                                       # gamma frozen (ESS_vec[3] = 0)
    assert r["par"][2] != st["par"][2]
    assert np.array_equal(r["OriginTraj"], st["OriginTraj"])
    assert r["logOrigin"] == pytest.approx(-0.5 * np.sum(st["OriginTraj"] ** 2), rel=1e-14)
    # the accepted state is self-consistent
    chk = orc.full_loglik(orc.Cfg(g.x_r, g.x_i, 42), orc.Init(g.init), r["par"][2:], r["par"][:2], g.times, 50,
                          r["OriginTraj"], g.t_correct)
    assert chk["coalLog"] == r["CoalLog"]


def test_chain_iterations_run_and_stay_finite(changepoint):
    g = changepoint
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    st = driver.init_state(g, par0, g.setting("control.traj"))
    # the vignette's chain starts at coalLog = l1[1] - (first updates); the initial value is deterministic
    assert np.isfinite(st["coalLog"]) and st["coalLog"] > -1e7
    opts = driver.opts_from_golden(g)
    rng = np.random.default_rng(3)
    for it in range(3):
        st = driver.eslice_iteration(st, g, opts, rng.standard_normal(43 + 80), rng.random(47))
        assert np.isfinite(st["coalLog"]) and st["coalLog"] > -1e7
        assert np.min(st["LatentTraj"][:, 1:]) >= 0
    assert st["n_full"] >= 3 * 3 and st["n_cheap"] >= 3
