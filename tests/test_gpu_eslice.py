"""GPU parity of the Metropolis / elliptical-slice operators and of the resident chain driver against
the CPU oracle on identical pre-drawn normals / uniforms: same accept / shrink decision sequences,
values within 1e-10 relative (BASELINE.json north_star: "ESlice chains step-for-step equal")."""
import numpy as np
import pytest

from conftest import relerr
from oracle import driver
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    import lnaphylodyn_b200 as m
    from lnaphylodyn_b200 import _lib
    assert _lib.lib().lna_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return m.api


def state_of(g):
    return driver.init_state(g, g.final["par"].ravel(), g.final["OriginTraj"])


def test_param_slice_updates(api, changepoint):
    g = changepoint
    par = g.final["par"].ravel()
    rng = np.random.default_rng(0)
    opts = driver.opts_from_golden(g)
    for th in (0.3, -1.7, 5.9):
        nu = rng.standard_normal(43)
        for ess in ([0, 1, 0, 0, 1, 1, 0], [1, 1, 1, 0, 1, 1, 0], [0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0, 0]):
            o = orc.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, opts["priorList"])
            d = api.Param_Slice_update_all2(par, g.x_r, g.x_i, th, nu, ess, opts["priorList"])
            assert relerr(d, o) < 1e-13
        nc = rng.standard_normal(39) / par[43]
        assert relerr(api.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nc), orc.Param_Slice_update(par[2:], g.x_r, g.x_i, th, nc)) < 1e-13


@pytest.mark.parametrize("width", [0, 1, 4, 32])
def test_eslice_par_general_step_for_step(api, changepoint, width):
    g = changepoint
    st = state_of(g)
    opts = driver.opts_from_golden(g)
    B = 48
    rng = np.random.default_rng(100)
    nrm, unf = rng.standard_normal((B, 43)), rng.random((B, 22))
    unf[:6, 0] = 1.0 - 1e-9 * rng.random(6)      # log u ~ 0: forces long shrink sequences / the 20-cap
    d = api.ESlice_par_General(np.tile(st["par"], (B, 1)), g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i,
                               g.init, g.gridsize, opts["ESS_vec"], st["coalLog"], g.t_correct, normals=nrm,
                               uniforms=unf, spec_width=width)
    caps = 0
    for b in range(B):
        o = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                                   g.gridsize, opts["ESS_vec"], st["coalLog"], g.t_correct, normals=nrm[b],
                                   uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"], (b, d["n_evals"][b], o["n_evals"])
        assert (d["status"][b] & 8) == (o["status"] & 8)
        caps += int(bool(o["status"] & 8))
        assert relerr(d["par"][b], o["par"]) < TOL
        assert d["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)
        assert d["logOrigin"][b] == pytest.approx(o["logOrigin"], rel=1e-13)
        assert relerr(d["LatentTraj"][b][:, 1:], o["LatentTraj"][:, 1:]) < 1e-9
        assert relerr(d["betaN"][b], o["betaN"]) < TOL
        assert relerr(d["OdeTraj"][b], o["OdeTraj"]) < TOL
        assert relerr(d["FT"]["A"][b], o["FT"]["A"]) < 1e-9
    assert d["n_evals"].max() > 3  # the sample exercises real shrinking


@pytest.mark.parametrize("ess", [[0, 1, 0, 0, 1, 1, 1], [1, 1, 1, 0, 1, 1, 1]])
def test_eslice_par_general_self_mixing_branch(api, changepoint, ess):
    """ESS_vec[7] = 1 (the default ESS_vec of the R drivers): k*p extra normals are drawn and discarded and every candidate
    whose New_Param_List succeeds replaces OriginTraj_new by OriginTraj cos(theta) + OriginTraj_new sin(theta)
    (LNA_functional.cpp:1576-1578, 1660-1662) -- a running multiple of OriginTraj.  Step for step against the oracle (which
    follows the reference code as written here: tests/test_ref_vs_oracle.py), including the returned OriginTraj / logOrigin."""
    g = changepoint
    st = state_of(g)
    opts = driver.opts_from_golden(g)
    B = 24
    rng = np.random.default_rng(111)
    nrm, unf = rng.standard_normal((B, 43)), rng.random((B, 22))
    unf[:3, 0] = 1.0 - 1e-9 * rng.random(3)   # long shrink sequences
    d = api.ESlice_par_General(np.tile(st["par"], (B, 1)), g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i,
                               g.init, g.gridsize, ess, st["coalLog"], g.t_correct, normals=nrm, uniforms=unf)
    mixed = 0
    for b in range(B):
        o = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                                   g.gridsize, ess, st["coalLog"], g.t_correct, normals=nrm[b], uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"] and (d["status"][b] & 8) == (o["status"] & 8), b
        assert relerr(d["par"][b], o["par"]) < TOL and d["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)
        assert np.max(np.abs(d["OriginTraj"][b] - o["OriginTraj"])) < 1e-13 * np.max(np.abs(o["OriginTraj"]))
        assert d["logOrigin"][b] == pytest.approx(o["logOrigin"], rel=1e-12)
        assert relerr(d["LatentTraj"][b][:, 1:], o["LatentTraj"][:, 1:]) < 1e-9
        mixed += not np.array_equal(o["OriginTraj"], st["OriginTraj"])
    assert mixed >= B - 4 and d["n_evals"].max() > 3   # (with the increments rescaled even log u ~ 0 is beaten before the cap)


def test_chains_with_default_ess_vec_self_mixing(api, changepoint):
    """The R drivers' DEFAULT ESS_vec = (1,1,1,1,1,1,1): I0, R0, gamma, hyper and the change ratios rotate, and the
    self-mixing branch evaluates the candidates on a multiple of OriginTraj -- which update_Par_ESlice_combine then does
    NOT store (R/LNA_chpt_slice.R:456-472), so the chain keeps its OriginTraj.  Chains against the restated iteration."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B, iters, seed = 8, 4, 5
    ess = [1, 1, 1, 1, 1, 1, 1]
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    eps0 = 0.2 * np.random.default_rng(2).standard_normal((40, 2))
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), eps0)
    ref = [driver.init_state(g, par0, eps0) for _ in range(B)]
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    for it in range(iters):
        traj = it >= 2
        e = list(ess)
        if not traj:
            e[4] = 0
        nrm, unf = chm.draw_streams(B, ch.n_norm, seed, it)
        ch.step(chm.make_opts(prior, proposal, e, update_traj=traj), nrm, unf)
        s = ch.get()
        o = dict(driver.opts_from_golden(g, update_traj=traj), ESS_vec=e)
        for c in range(B):
            ref[c] = driver.eslice_iteration(ref[c], g, o, nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL, (it, c)
            assert s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL), (it, c)
            assert np.max(np.abs(s["OriginTraj"][c] - ref[c]["OriginTraj"])) < 1e-11
            assert s["n_full_evals"][c] == ref[c]["n_full"], (it, c)
    assert not np.array_equal(s["par"][:, 1], np.full(B, 1.0))   # I0 moves under ESS_vec[1]


def test_eslice_general_nc_kingman_branch(api, changepoint):
    """ESlice_general_NC(volz = FALSE): coal_loglik3 of LogTraj(newTraj) -- log(log I), NaN as soon as I < 1, and a NaN ends
    the loop and is returned; the 20-cap recomputes with volz_loglik_nh2 (:1717-1760).  Step for step against the oracle,
    which follows the reference code as written on the same cases (tests/test_ref_vs_oracle.py)."""
    g = changepoint
    rng0 = np.random.default_rng(9)
    seen_nan = seen_cap = seen_shrink = 0
    for s in range(14):
        par = g.final["par"].ravel().copy()
        par[2] *= np.exp(0.02 * (s % 3))
        par[1] = 0.5 if s >= 10 else 80.0
        eps = g.final["OriginTraj"] + 0.05 * rng0.standard_normal((40, 2))
        st = driver.init_state(g, par, eps)
        rng = np.random.default_rng([23, s])
        kw = dict(normals=rng.standard_normal(80), uniforms=rng.uniform(size=22))
        lam = (1.0, 300.0)[s % 2]
        lat0 = np.array(st["LatentTraj"], order="F")
        with np.errstate(invalid="ignore", divide="ignore"):
            base = orc.coal_loglik3(g.init, np.column_stack([lat0[:, 0], np.log(lat0[:, 1:])]), g.t_correct, lam, 1)
        cl = (base - 3.0, base + 0.5, 1e9, base)[s % 4] if np.isfinite(base) else -2000.0
        args = (eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, lam, cl, g.gridsize, False, "SIR")
        a, b = orc.ESlice_general_NC(*args, **kw), api.ESlice_general_NC(*args, **kw)
        assert a["n_evals"] == b["n_evals"], s
        assert np.allclose(a["OriginTraj"], b["OriginTraj"], rtol=1e-12, atol=1e-300)
        assert relerr(b["LatentTraj"][:, 1:], a["LatentTraj"][:, 1:]) < 1e-9
        if np.isnan(a["CoalLog"]):
            assert np.isnan(b["CoalLog"]) and (b["status"] & 4)
            seen_nan += 1
        else:
            assert b["CoalLog"] == pytest.approx(a["CoalLog"], rel=TOL)
        assert b["logOrigin"] == pytest.approx(a["logOrigin"], rel=1e-12)
        seen_cap += bool(a["status"] & 8)
        seen_shrink += a["n_evals"] > 2
    assert seen_nan >= 1 and seen_cap >= 1 and seen_shrink >= 2
    from lnaphylodyn_b200._lib import LnaError
    with pytest.raises(LnaError, match="99999999"):
        api.ESlice_general_NC(eps, st["Ode"], st["FT"], par[:2], g.init, st["betaN"], g.t_correct, 1.0, -99999999.0,
                              g.gridsize, False, "SIR", **kw)


def test_eslice_par_general_two_phase_window(api, changepoint):
    """B inside the window where the automatic policy runs two compacted phases (candidates 1..8 for
    every chain, then 9..21 sixteen-wide for the unresolved ones): same chains as the sequential loop.
    Spot-checked against the oracle, fully checked against the single-phase GPU path."""
    g = changepoint
    st = state_of(g)
    opts = driver.opts_from_golden(g)
    B = 3000
    rng = np.random.default_rng(101)
    nrm, unf = rng.standard_normal((B, 43)), rng.random((B, 22))
    unf[:40, 0] = 1.0 - 1e-9 * rng.random(40)
    call = lambda w: api.ESlice_par_General(np.tile(st["par"], (B, 1)), g.times, st["OriginTraj"], opts["priorList"],
                                            g.x_r, g.x_i, g.init, g.gridsize, opts["ESS_vec"], st["coalLog"],
                                            g.t_correct, normals=nrm, uniforms=unf, spec_width=w)
    d2, d1 = call(0), call(4)
    assert np.array_equal(d2["n_evals"], d1["n_evals"]) and np.array_equal(d2["status"], d1["status"])
    assert np.array_equal(d2["par"], d1["par"]) and np.array_equal(d2["CoalLog"], d1["CoalLog"])
    assert np.array_equal(d2["LatentTraj"], d1["LatentTraj"]) and np.array_equal(d2["FT"]["A"], d1["FT"]["A"])
    assert (d2["n_evals"] > 8).sum() > 100 and (d2["status"] & 8).sum() >= 1
    for b in list(range(0, 40, 4)) + list(range(40, B, 300)):
        o = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                                   g.gridsize, opts["ESS_vec"], st["coalLog"], g.t_correct, normals=nrm[b],
                                   uniforms=unf[b])
        assert d2["n_evals"][b] == o["n_evals"] and relerr(d2["par"][b], o["par"]) < TOL
        assert d2["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)


@pytest.mark.parametrize("plan", [-8, -4, -2, -1])
def test_eslice_phase_plans_equal_sequential(api, changepoint, plan):
    """Every automatic phase plan (compaction between phases, several in-kernel waves per phase) must give the
    chains of the plain one-candidate-at-a-time evaluation, bit for bit."""
    g = changepoint
    st = state_of(g)
    opts = driver.opts_from_golden(g)
    B = 700
    rng = np.random.default_rng(303)
    nrm, unf = rng.standard_normal((B, 43)), rng.random((B, 22))
    unf[:20, 0] = 1.0 - 1e-9 * rng.random(20)
    call = lambda w: api.ESlice_par_General(np.tile(st["par"], (B, 1)), g.times, st["OriginTraj"], opts["priorList"],
                                            g.x_r, g.x_i, g.init, g.gridsize, opts["ESS_vec"], st["coalLog"],
                                            g.t_correct, normals=nrm, uniforms=unf, spec_width=w)
    d, ref = call(plan), call(1)
    for k in ("n_evals", "status", "par", "CoalLog", "LatentTraj", "betaN"):
        assert np.array_equal(d[k], ref[k]), k
    assert np.array_equal(d["FT"]["A"], ref["FT"]["A"]) and (ref["status"] & 8).sum() >= 1


def test_eslice_general_nc_step_for_step(api, changepoint):
    g = changepoint
    st = state_of(g)
    B = 64
    rng = np.random.default_rng(200)
    nrm, unf = rng.standard_normal((B, 80)), rng.random((B, 22))
    unf[:4, 0] = 1.0 - 1e-12
    rep = lambda a: np.tile(a, (B,) + (1,) * np.ndim(a))
    d = api.ESlice_general_NC(rep(st["OriginTraj"]), rep(st["Ode"]), {"A": rep(st["FT"]["A"]), "Sigma": rep(st["FT"]["Sigma"])},
                              st["par"][:2], g.init, rep(st["betaN"]), g.t_correct, 1.0, st["coalLog"], g.gridsize, True,
                              "SIR", normals=nrm, uniforms=unf)
    for b in range(B):
        o = orc.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, st["betaN"], g.t_correct,
                                  1.0, st["coalLog"], g.gridsize, True, "SIR", normals=nrm[b], uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"], (b, d["n_evals"][b], o["n_evals"])
        assert (d["status"][b] & 8) == (o["status"] & 8)
        assert d["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)
        assert d["logOrigin"][b] == pytest.approx(o["logOrigin"], rel=1e-12)
        assert np.max(np.abs(d["OriginTraj"][b] - o["OriginTraj"])) < 1e-12
        assert relerr(d["LatentTraj"][b][:, 1:], o["LatentTraj"][:, 1:]) < 1e-9
    assert d["n_evals"].max() > 2
    # coal_log = -99999999 recomputes the current likelihood first
    d2 = api.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, st["betaN"], g.t_correct,
                               1.0, -99999999.0, g.gridsize, True, "SIR", normals=nrm[9], uniforms=unf[9])
    o2 = orc.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, st["betaN"], g.t_correct,
                               1.0, -99999999.0, g.gridsize, True, "SIR", normals=nrm[9], uniforms=unf[9])
    assert d2["n_evals"] == o2["n_evals"] and d2["CoalLog"] == pytest.approx(o2["CoalLog"], rel=TOL)


@pytest.mark.parametrize("width", [0, 8])
def test_eslice_change_points_step_for_step(api, constant, width):
    """General_MCMC2's change-point update (constant_sim vignette)."""
    g = constant
    st = driver.init_state(g, g.final["par"].ravel(), g.final["OriginTraj"])
    B = 32
    rng = np.random.default_rng(300)
    nrm, unf = rng.standard_normal((B, 39)), rng.random((B, 22))
    unf[:4, 0] = 1.0 - 1e-10
    d = api.ESlice_change_points(np.tile(st["par"][2:], (B, 1)), st["par"][:2], g.times, st["OriginTraj"], g.x_r, g.x_i,
                                 g.init, g.gridsize, st["coalLog"], g.t_correct, normals=nrm, uniforms=unf,
                                 spec_width=width)
    for b in range(B):
        o = orc.ESlice_change_points(st["par"][2:], st["par"][:2], g.times, st["OriginTraj"], g.x_r, g.x_i, g.init,
                                     g.gridsize, st["coalLog"], g.t_correct, normals=nrm[b], uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"], (b, d["n_evals"][b], o["n_evals"])
        assert relerr(d["param"][b], o["param"]) < TOL
        assert d["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)
        assert d["LogChProb"][b] == pytest.approx(o["LogChProb"], rel=1e-9, abs=1e-12)
        assert relerr(d["LatentTraj"][b][:, 1:], o["LatentTraj"][:, 1:]) < 1e-9


@pytest.mark.parametrize("mh_pair", ["1", "0"])
def test_chains_step_for_step(api, changepoint, monkeypatch, mh_pair):
    """B chains x 6 iterations of the General_MCMC_with_ESlice body (2 burn-in style + 4 full) from the
    vignette's initial state: every chain must reproduce the oracle's chain step for step -- with the two Metropolis
    entries evaluated speculatively in one launch (small batches) and one after the other (batches that fill the machine)."""
    from lnaphylodyn_b200 import chains as chm
    monkeypatch.setenv("LNA_MH_PAIR", mh_pair)
    g = changepoint
    B, iters, seed = 12, 6, 77
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    ref = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]
    s0 = ch.get()
    assert s0["coalLog"][0] == pytest.approx(ref[0]["coalLog"], rel=TOL)
    for it in range(iters):
        traj = it >= 2
        nrm, unf = chm.draw_streams(B, ch.n_norm, seed, it)
        ch.step(chm.vignette_opts(g, update_traj=traj), nrm, unf)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.eslice_iteration(ref[c], g, driver.opts_from_golden(g, update_traj=traj), nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL, (it, c)
            assert s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL), (it, c)
            assert np.max(np.abs(s["OriginTraj"][c] - ref[c]["OriginTraj"])) < 1e-11
            assert relerr(s["LatentTraj"][c][:, 1:], ref[c]["LatentTraj"][:, 1:]) < 1e-9
            assert s["logOrigin"][c] == pytest.approx(ref[c]["logOrigin"], rel=1e-11, abs=1e-300)
            assert s["n_full_evals"][c] == ref[c]["n_full"], (it, c)
            assert s["n_cheap_evals"][c] == ref[c]["n_cheap"], (it, c)
            assert s["n_mh_accept"][c] == ref[c]["n_mh_accept"], (it, c)
    assert np.all(np.isfinite(s["coalLog"]))


def test_sharded_chains_equal_single_device(api, changepoint):
    """chains are independent, so a batch sharded over two GPUs must reproduce the
    single-GPU batch per chain, bit for bit."""
    from lnaphylodyn_b200 import _lib, chains as chm
    if _lib.lib().lna_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = changepoint
    B, iters, seed = 16, 3, 9
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    single = chm.Chains(api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init, device=0),
                        np.tile(par0, (B, 1)), g.setting("control.traj"))
    shards = []
    for r in range(2):
        b, e = chm.shard_range(B, r, 2)
        pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init, device=r)
        shards.append((b, e, chm.Chains(pb, np.tile(par0, (e - b, 1)), g.setting("control.traj"))))
    opts = chm.vignette_opts(g, update_traj=True)
    for it in range(iters):
        nrm, unf = chm.draw_streams(B, single.n_norm, seed, it)
        single.step(opts, nrm, unf)
        for b, e, ch in shards:
            ch.step(opts, nrm[b:e], unf[b:e])
    s = single.get()
    for b, e, ch in shards:
        t = ch.get()
        for k in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals"):
            assert np.array_equal(t[k], s[k][b:e]), k


def test_chains_mcmc2_step_for_step(api, constant):
    """General_MCMC2 (constant_sim vignette): Update_Param_each x3, update_hyper, change-point ESS,
    trajectory ESS -- B chains x 5 iterations from the vignette's initial state against the oracle."""
    from lnaphylodyn_b200 import chains as chm
    g = constant
    B, iters, seed = 10, 5, 41
    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                           g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    ref = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    accepted = 0
    for it in range(iters):
        burn = it < 1
        rng = np.random.default_rng([seed, it])
        nrm = rng.standard_normal((B, 39 + 80))
        unf = rng.random((B, chm.ITER2_UNIF))
        ch.step_mcmc2(chm.make_mcmc2_opts(prior, proposal, update_hyper=not burn, update_traj=not burn), nrm, unf)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.mcmc2_iteration(ref[c], g, driver.mcmc2_opts_from_golden(g, burn=burn), nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL, (it, c)
            assert s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL), (it, c)
            assert np.max(np.abs(s["OriginTraj"][c] - ref[c]["OriginTraj"])) < 1e-11
            assert s["n_full_evals"][c] == ref[c]["n_full"] and s["n_cheap_evals"][c] == ref[c]["n_cheap"]
            assert s["n_mh_accept"][c] == ref[c]["n_mh_accept"], (it, c)
        accepted = int(s["n_mh_accept"].sum())
    assert 0 < accepted < B * iters * 4


def test_device_rng_streams_and_replay(api, changepoint):
    """On-device Philox streams: N(0,1)/U(0,1) moments, independence of the sharding (chain_offset), and a
    chain advanced with step_rng replays exactly through the oracle on the streams read back to the host."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B = 8
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    big = chm.Chains(pb, np.tile(par0, (2048, 1)), g.setting("control.traj"))
    n, u = big.device_streams(seed=7, iteration=3)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01 and abs(u.mean() - 0.5) < 0.005
    assert 0 < u.min() and u.max() < 1 and abs(np.mean(n ** 4) - 3) < 0.1
    assert abs(np.corrcoef(n[:, 0], n[:, 1])[0, 1]) < 0.08 and len(np.unique(u)) == u.size
    n2, u2 = ch.device_streams(seed=7, iteration=3, chain_offset=1000)
    assert np.array_equal(n2, n[1000:1008]) and np.array_equal(u2, u[1000:1008])   # sharding-independent
    ref = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]
    for it in range(3):
        opts = chm.vignette_opts(g, update_traj=it >= 1)
        nrm, unf = ch.device_streams(seed=11, iteration=it, chain_offset=40)
        ch.step_rng(opts, seed=11, iteration=it, chain_offset=40)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.eslice_iteration(ref[c], g, driver.opts_from_golden(g, update_traj=it >= 1), nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL and s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL)
            assert s["n_full_evals"][c] == ref[c]["n_full"]


def test_device_rng_streams_do_not_alias_across_seeds(api, changepoint):
    """Seed and chain index occupy different Philox words: (seed 0, chain 1) and (seed 1, chain 0) -- and every other pair
    with seed ^ chain equal -- draw different streams, so replicate runs with seeds 0, 1, 2, ... are independent."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (8, 1)), g.setting("control.traj"))
    streams = {s: ch.device_streams(seed=s, iteration=5) for s in range(4)}
    rows = {}
    for s, (n, u) in streams.items():
        for c in range(8):
            rows[(s, c)] = np.concatenate([n[c], u[c]])
    keys = list(rows)
    for i, a in enumerate(keys):
        for b in keys[i + 1:]:
            assert not np.any(rows[a] == rows[b]), (a, b)   # no shared value at any position
    # iteration is a separate word too
    n5, _ = ch.device_streams(seed=0, iteration=5)
    n6, _ = ch.device_streams(seed=0, iteration=6)
    assert not np.any(n5 == n6)
    u = np.concatenate([streams[s][1].ravel() for s in streams])
    assert u.min() > 0.0 and u.max() < 1.0


def test_initial_log_origin_is_recorded(api, changepoint):
    """MCMC_obj$logOrigin starts at -sum(OriginTraj^2)/2 (R/general_change_point.R:530) and update_Par_ESlice_combine does
    not touch it, so iterations before the first trajectory slice update record that value in `l`."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B = 6
    rng = np.random.default_rng(5)
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    eps = 0.3 * rng.standard_normal((B, 40, 2))
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), eps)
    want = -0.5 * np.sum(eps ** 2, axis=(1, 2))
    assert np.allclose(ch.get()["logOrigin"], want, rtol=1e-13, atol=0)
    ref = [driver.init_state(g, par0, eps[c]) for c in range(B)]
    for it in range(2):   # burn-in style iterations: no trajectory update
        nrm, unf = chm.draw_streams(B, ch.n_norm, 3, it)
        ch.step(chm.vignette_opts(g, update_traj=False), nrm, unf)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.eslice_iteration(ref[c], g, driver.opts_from_golden(g, update_traj=False), nrm[c], unf[c])
            assert s["logOrigin"][c] == pytest.approx(ref[c]["logOrigin"], rel=1e-13) == pytest.approx(want[c], rel=1e-13)
            assert relerr(s["par"][c], ref[c]["par"]) < TOL


def test_chain_stream_groups_equal_single_group(api, changepoint, monkeypatch):
    """Large batches run as several stream groups (overlap of latency-bound and throughput-bound launches), optionally
    free-running (not joined after every step) and with the iterations replayed from CUDA graphs: per chain the result must
    not depend on the grouping, the joining or the replay, and must match the oracle."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B, iters = 2101, 5
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    monkeypatch.setenv("LNA_CHAIN_GROUPS", "3")
    grouped = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=True)
    assert grouped.groups == 3
    monkeypatch.setenv("LNA_CHAIN_GROUPS", "1")
    single = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    monkeypatch.delenv("LNA_CHAIN_GROUPS")
    ref = {c: driver.init_state(g, par0, g.setting("control.traj")) for c in (0, 699, 700, 701, 1400, 2100)}
    for it in range(iters):
        opts = chm.vignette_opts(g, update_traj=it >= 1)
        nrm, unf = grouped.device_streams(seed=3, iteration=it)
        grouped.step_rng(opts, seed=3, iteration=it)
        single.step(opts, nrm, unf)
        for c in ref:
            ref[c] = driver.eslice_iteration(ref[c], g, driver.opts_from_golden(g, update_traj=it >= 1), nrm[c], unf[c])
    a, b = grouped.get(), single.get()
    for k in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals", "n_mh_accept"):
        assert np.array_equal(a[k], b[k]), k
    for c, r in ref.items():
        assert relerr(a["par"][c], r["par"]) < TOL and a["coalLog"][c] == pytest.approx(r["coalLog"], rel=TOL)
        assert a["n_full_evals"][c] == r["n_full"]


@pytest.mark.parametrize("B", [5, 96, 700])
def test_three_warp_pipeline_is_bit_identical(api, changepoint, monkeypatch, B):
    """csrc/lna_split.cuh: the latency-bound launches evaluate every candidate with a pipeline of three warps (mean ODE ->
    LNA moments -> interval epilogues) instead of one.  Same expressions in the same order, so the chains it produces must
    equal the one-warp kernels' BIT FOR BIT -- parameters, latent trajectories, likelihoods and counters -- over several
    iterations, for ragged batches (idle lanes), partially filled warps and multi-wave slice launches."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    iters, seed = 5, 123
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LNA_SPLIT", mode)
        ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
        for it in range(iters):
            ch.step_rng(chm.vignette_opts(g, update_traj=it >= 1), seed, it)
        runs[mode] = ch.get()
    a, b = runs["0"], runs["1"]
    for key in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals", "n_mh_accept"):
        assert np.array_equal(a[key], b[key]), key
    assert np.all(np.isfinite(b["coalLog"])) and len(np.unique(b["coalLog"])) > B // 2


def test_three_warp_pipeline_standalone_entry_points(api, changepoint, monkeypatch):
    """The same through the standalone operators (ESlice_par_General with several in-kernel waves, ESlice_change_points,
    Update_Param), including a proposal whose Cholesky fails and the 20-shrink cap."""
    g = changepoint
    par = g.final["par"].ravel().copy()
    eps = g.final["OriginTraj"].copy()
    rng = np.random.default_rng(5)
    cl = float(g.final["coalLog"][0])
    PRI = [(1.0, 1.0), (0.7, 0.5), (-1.7, 0.1), (3.0, 0.2)]
    wild = [(1.0, 1.0), (0.7, 40.0), (-1.7, 0.1), (3.0, 0.2)]
    cases = []
    for s in range(6):
        nrm, unf = rng.standard_normal(43), rng.uniform(size=22)
        cases.append(("par", PRI, cl, nrm, unf))
        cases.append(("par", wild, cl, 3.0 * nrm, unf))
    cases.append(("par", PRI, 1e6, rng.standard_normal(43), rng.uniform(size=22)))  # cap
    nrm_cp, unf_cp = rng.standard_normal(39), rng.uniform(size=22)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LNA_SPLIT", mode)
        res = []
        for kind, pri, c0, nrm, unf in cases:
            for W in (0, 4):
                r = api.ESlice_par_General(par, g.times, eps, pri, g.x_r, g.x_i, g.init, g.gridsize, [0, 1, 0, 0, 1, 1, 0], c0,
                                           g.t_correct, normals=nrm, uniforms=unf, spec_width=W)
                res.append((r["par"], r["LatentTraj"], r["CoalLog"], r["n_evals"], r["status"], r["FT"]["Sigma"]))
        r = api.ESlice_change_points(par[2:], par[:2], g.times, eps, g.x_r, g.x_i, g.init, g.gridsize, cl, g.t_correct,
                                     normals=nrm_cp, uniforms=unf_cp)
        res.append((r["param"], r["LatentTraj"], r["CoalLog"], r["n_evals"], r["status"], r["FT"]["A"]))
        out[mode] = res
    for x, y in zip(out["0"], out["1"]):
        for u, v in zip(x, y):
            assert np.array_equal(np.asarray(u), np.asarray(v))


def test_chains_mcmc2_joint_stage_step_for_step(api, constant):
    """General_MCMC2's third stage (i >= warmup2): update_Param_joint(method = "admcmc") over (I0, R0, gamma, hyper) with
    per-chain proposal transforms, then change-point and trajectory slice updates -- B chains x 5 iterations against the
    restated R code (oracle/driver.py::mcmc2_joint_iteration) on the same streams and the same transforms."""
    from lnaphylodyn_b200 import chains as chm
    g = constant
    B, iters, seed = 9, 5, 53
    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                           g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    ref = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    entries = ((1, 1, 1), (2, 0, 2), (3, 1, 3), (4, 0, 5))  # (joint index, isLog, priorId): I0, R0, gamma, hyper
    rng = np.random.default_rng(seed)
    # a different (correlated) proposal covariance per chain, small enough for a healthy acceptance rate
    A = rng.standard_normal((B, 4, 4)) * np.array([0.02, 0.004, 0.01, 0.3])[None, :, None]
    pcov = A @ np.transpose(A, (0, 2, 1))
    T = chm.mvrnorm_transform(pcov)
    assert np.allclose(T @ np.transpose(T, (0, 2, 1)), pcov, rtol=1e-9, atol=1e-18)
    ch.set_joint_transform(T)
    opts = chm.make_joint_opts(prior, entries)
    oopts = dict(driver.mcmc2_opts_from_golden(g), joint=[(j, lg, 1 if pid == 2 else 2 if pid == 5 else 0, tuple(prior[pid - 1][:2]))
                                                          for j, lg, pid in entries])
    for it in range(iters):
        nrm = rng.standard_normal((B, 39 + 80 + 4))
        unf = rng.random((B, chm.ITER2_UNIF))
        ch.step_mcmc2(opts, nrm, unf)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.mcmc2_joint_iteration(ref[c], g, oopts, T[c], nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL, (it, c)
            assert s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL), (it, c)
            assert np.max(np.abs(s["OriginTraj"][c] - ref[c]["OriginTraj"])) < 1e-11
            assert s["n_full_evals"][c] == ref[c]["n_full"] and s["n_cheap_evals"][c] == ref[c]["n_cheap"]
            assert s["n_mh_accept"][c] == ref[c]["n_mh_accept"], (it, c)
    acc = int(s["n_mh_accept"].sum())
    assert 0 < acc < B * iters
    moved = np.abs(s["par"][:, [1, 2, 3, 43]] / par0[[1, 2, 3, 43]] - 1.0) > 0
    assert moved.any(axis=0).all()  # every jointly proposed parameter moved in some chain


def test_chains_general_entry_list_with_hyper_proposal(api, changepoint):
    """Update_Param_each_Norm with hyper = T (R/LNA_chpt_slice.R:219-241): the hyper entry proposes hyper' = hyper e^u under a
    log-normal prior and rescales the change ratios; plus an I0 entry on the log scale and an R0 entry on the raw scale --
    the general entry list against the restated R iteration, step for step."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B, iters, seed = 7, 5, 91
    nch = 39
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
    ref = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    entries = ((2, 1, 1), (3, 0, 2), (4, 1, 3), (nch + 5, 0, 5))  # I0 (log), R0 (raw), gamma (log), hyper
    ess = [0, 1, 0, 0, 1, 1, 0]
    opts = chm.make_opts_general(prior, proposal, ess, entries, hyper=True, update_traj=True)
    oopts = {"priorList": [prior[0][:2], prior[1][:2], prior[2][:2], prior[4][:2]], "ESS_vec": ess, "update_traj": True,
             "hyper": True,
             "entries": [(1, 1, tuple(prior[0][:2]), float(proposal[0][0])), (2, 0, tuple(prior[1][:2]), float(proposal[1][0])),
                         (3, 1, tuple(prior[2][:2]), float(proposal[2][0])), (-1, 0, tuple(prior[4][:2]), float(proposal[4][0]))]}
    for it in range(iters):
        rng = np.random.default_rng([seed, it])
        nrm, unf = rng.standard_normal((B, ch.n_norm)), rng.random((B, chm.ITER_UNIF_GENERAL))
        ch.step(opts, nrm, unf)
        s = ch.get()
        for c in range(B):
            ref[c] = driver.eslice_iteration_general(ref[c], g, oopts, nrm[c], unf[c])
            assert relerr(s["par"][c], ref[c]["par"]) < TOL, (it, c)
            assert s["coalLog"][c] == pytest.approx(ref[c]["coalLog"], rel=TOL), (it, c)
            assert s["n_full_evals"][c] == ref[c]["n_full"] and s["n_mh_accept"][c] == ref[c]["n_mh_accept"], (it, c)
    assert len(np.unique(s["par"][:, 43])) > 1 and len(np.unique(s["par"][:, 1])) > 1  # hyper and I0 moved


@pytest.mark.parametrize("B", [3, 40, 700, 1500, 3000])
def test_metropolis_tree_is_bit_identical(api, constant, monkeypatch, B):
    """General_MCMC2's four consecutive Metropolis entries (I0, R0, gamma, hyper) evaluated speculatively in one latency
    period -- candidate (entry e, accepted subset A) on lane 2^e - 1 + A, accept tests chained afterwards -- must give the
    chains the sequential entries give, bit for bit, counters included.  (Up to 1184 chains one tree of four entries; 1500
    chains: trees of 3 + 1; 3000 chains, run as two stream groups: two trees of 2.)"""
    from lnaphylodyn_b200 import chains as chm
    g = constant
    iters, seed = 6 if B < 1000 else 3, 77
    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                           g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LNA_MH_TREE", mode)
        ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
        for it in range(iters):
            rng = np.random.default_rng([seed, it])
            nrm, unf = rng.standard_normal((B, 39 + 80)), rng.random((B, chm.ITER2_UNIF))
            ch.step_mcmc2(chm.make_mcmc2_opts(prior, proposal, update_hyper=it >= 1, update_traj=it >= 1), nrm, unf)
        runs[mode] = ch.get()
    for key in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals", "n_mh_accept"):
        assert np.array_equal(runs["0"][key], runs["1"][key]), key
    acc = runs["1"]["n_mh_accept"]
    assert 0 < acc.sum() < B * iters * 4


@pytest.mark.parametrize("hyper", [False, True])
def test_metropolis_tree_general_entries_bit_identical(api, changepoint, monkeypatch, hyper):
    """The same for General_MCMC_with_ESlice's general Update_Param_each_Norm entry list (S0 and R0 walks + the hyper entry,
    no-op or log-normal walk): tree of 7 candidates on 8 lanes against the sequential entries."""
    from lnaphylodyn_b200 import chains as chm
    g = changepoint
    B, iters, seed, nch = 37, 5, 5, 39
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"),
                           g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    prior[0] = np.array([13.8, 0.5, 1.0, 1.0])  # a pop_pr under which a walk on log S0 is accepted now and then
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    entries = ((1, 1, 1), (3, 1, 2), (nch + 5, 0, 5))
    opts = chm.make_opts_general(prior, proposal, [0, 1, 0, 0, 1, 1, 0], entries, hyper=hyper, update_traj=True)
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LNA_MH_TREE", mode)
        ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
        for it in range(iters):
            rng = np.random.default_rng([seed, it])
            ch.step(opts, rng.standard_normal((B, ch.n_norm)), rng.random((B, chm.ITER_UNIF_GENERAL)))
        runs[mode] = ch.get()
    for key in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals", "n_mh_accept"):
        assert np.array_equal(runs["0"][key], runs["1"][key]), key
    assert len(np.unique(runs["1"]["par"][:, 0])) > 1  # S0 moved in some chains


def test_slice_kernel_variants_give_the_same_chains(changepoint):
    """The parameter slice update has three device schedules -- speculative waves (lna_eslice.cuh), barrier-free lanes with
    early rejection (lna_async.cuh), sequential candidates with early rejection, parked state and compaction (lna_seg.cuh).
    Whatever the schedule, the accepted candidate is the one the sequential rule picks: 8 iterations of 96 chains end in the
    same state (decisions identical; values to the last bits of the per-interval dt operands)."""
    import os
    from lnaphylodyn_b200 import api, chains as chm
    g = changepoint
    B = 96
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    keys = ("LNA_ESS_ASYNC", "LNA_ESS_ASYNC_W", "LNA_ESS_ASYNC_WPS", "LNA_ESS_SEG", "LNA_ESS_SEG_T", "LNA_GRAPH")
    saved = {k: os.environ.get(k) for k in keys}
    # second batch: more chains than a persistent launch of one 32-lane group per warp and one warp per SM sub-partition holds
    # (592 on a B200), so that groups REFILL from the batch counter; 16 lanes per chain is the default from 1007 chains on
    cases = ((96, (("waves", {"LNA_ESS_ASYNC": "0"}), ("barrier-free", {"LNA_ESS_ASYNC_W": "8"}),
                   ("barrier-free-4", {"LNA_ESS_ASYNC_W": "4"}), ("barrier-free-16", {"LNA_ESS_ASYNC_W": "16"}),
                   ("sequential", {"LNA_ESS_SEG": "1", "LNA_ESS_SEG_T": "7"}))),
             (1100, (("waves", {"LNA_ESS_ASYNC": "0"}), ("default", {}),
                     ("barrier-free-32-refill", {"LNA_ESS_ASYNC_W": "32", "LNA_ESS_ASYNC_WPS": "1"}),
                     ("barrier-free-8-refill", {"LNA_ESS_ASYNC_W": "8", "LNA_ESS_ASYNC_WPS": "1"}))))
    try:
        for B, variants in cases:
            results = {}
            for name, env in variants:
                for k in keys:
                    os.environ.pop(k, None)
                os.environ.update(env)
                os.environ["LNA_GRAPH"] = "0"  # (a captured graph would freeze the first variant's launches)
                ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"))
                for it in range(8 if B < 1000 else 5):
                    ch.step_rng(chm.vignette_opts(g, update_traj=it >= 2), 99, it)
                results[name] = ch.get()
            ref = results["waves"]
            for name, r in results.items():
                assert np.array_equal(r["n_full_evals"], ref["n_full_evals"]) and np.array_equal(r["n_mh_accept"], ref["n_mh_accept"]), (B, name)
                assert relerr(r["par"], ref["par"]) < 1e-12 and relerr(r["coalLog"], ref["coalLog"]) < 1e-12, (B, name)
                assert relerr(r["LatentTraj"], ref["LatentTraj"]) < 1e-10 and np.array_equal(r["status"], ref["status"]), (B, name)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
