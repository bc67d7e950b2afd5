"""Edge cases of the GPU path against the oracle: no change points (constant R0), change points off the
coarse grid, other grid sizes (k > 32 rows per warp pass, k > 64 = shared-memory dt table path),
single-item and empty batches, non-uniform tail of the time grid."""
import numpy as np
import pytest

from conftest import relerr
from lnaphylodyn_b200 import genealogy as gen
from oracle import driver
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    import lnaphylodyn_b200 as m
    from lnaphylodyn_b200 import _lib
    assert _lib.lib().lna_device_count() > 0
    return m.api


def make_init(g, times, gridsize):
    a = {k: g.setting("Init.args." + k) for k in ("samp_times", "n_sampled", "coal_times")}
    return gen.coal_lik_init(a["samp_times"], a["n_sampled"], a["coal_times"], times[::gridsize])


def compare(api, times, x_r, x_i, init, gridsize, t_correct, param, initial, eps):
    pb = api.problem_for(times, x_r, x_i, gridsize, t_correct, init)
    res = pb.full_loglik(param, initial, eps, want=("coalLog", "FT", "Ode", "betaN", "LatentTraj"))
    cfg, ini = orc.Cfg(x_r, x_i, param.shape[1]), orc.Init(init)
    nok = 0
    for b in range(param.shape[0]):
        o = orc.full_loglik(cfg, ini, param[b], initial[b], times, gridsize, eps[b], t_correct)
        assert o["status"] == res["status"][b]
        assert relerr(res["Ode"][b], o["Ode"]) < TOL and relerr(res["betaN"][b], o["betaN"]) < TOL
        assert relerr(res["FT"]["A"][b], o["FT"]["A"]) < 1e-9
        assert np.max(np.abs(res["LatentTraj"][b][:, 1:] - o["LatentTraj"][:, 1:]) /
                      np.maximum(np.abs(o["LatentTraj"][:, 1:]), 1.0)) < TOL
        if o["coalLog"] == orc.SENTINEL:
            assert res["coalLog"][b] == orc.SENTINEL
        else:
            nok += 1
            assert res["coalLog"][b] == pytest.approx(o["coalLog"], rel=TOL)
    assert nok >= param.shape[0] // 4
    return pb, res


def draws(rng, B, nch, k, eps_scale=0.2):
    param = np.empty((B, 2 + nch + 1))
    param[:, 0] = 2.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 1] = 0.2 * np.exp(0.05 * rng.standard_normal(B))
    param[:, 2:2 + nch] = np.exp(0.03 * rng.standard_normal((B, nch)))
    param[:, -1] = 20.0
    eps = eps_scale * rng.standard_normal((B, k, 2))
    eps[:, : k // 3] = np.abs(eps[:, : k // 3])
    return param, np.tile([1e6, 1.0], (B, 1)), eps


def test_constant_R0_no_change_points(api, changepoint):
    """BASELINE.json: 'the constant-R0 SIR problem' -- x_i = (0, 2, 0, 1), param = (R0, gamma, hyper)."""
    g = changepoint
    rng = np.random.default_rng(1)
    param, initial, eps = draws(rng, 24, 0, 40)
    x_r, x_i = [1e6], [0, 2, 0, 1]
    compare(api, g.times, x_r, x_i, g.init, 50, g.t_correct, param, initial, eps)
    # parameter ESS without change points (rotates R0 and hyper only)
    par = np.concatenate([[1e6, 1.0], param[0]])
    st = orc.full_loglik(orc.Cfg(x_r, x_i, 3), orc.Init(g.init), par[2:], par[:2], g.times, 50, eps[0], g.t_correct)
    opts = driver.opts_from_golden(g)
    B = 16
    nrm, unf = rng.standard_normal((B, 4)), rng.random((B, 22))
    d = api.ESlice_par_General(np.tile(par, (B, 1)), g.times, eps[0], opts["priorList"], x_r, x_i, g.init, 50,
                               opts["ESS_vec"], st["coalLog"], g.t_correct, normals=nrm, uniforms=unf)
    for b in range(B):
        o = orc.ESlice_par_General(par, g.times, eps[0], opts["priorList"], x_r, x_i, g.init, 50, opts["ESS_vec"],
                                   st["coalLog"], g.t_correct, normals=nrm[b], uniforms=unf[b])
        assert d["n_evals"][b] == o["n_evals"] and relerr(d["par"][b], o["par"]) < TOL
        assert d["CoalLog"][b] == pytest.approx(o["CoalLog"], rel=TOL)


def test_change_points_off_the_coarse_grid(api, changepoint):
    g = changepoint
    rng = np.random.default_rng(2)
    cht = np.array([0.0, 3.33, 13.37, 13.38, 55.5, 55.55, 77.025, 99.99, 100.0, 250.0])
    param, initial, eps = draws(rng, 24, len(cht), 40)
    param[:, 2:2 + len(cht)] = np.exp(0.3 * rng.standard_normal((24, len(cht))))
    compare(api, g.times, np.concatenate([[1e6], cht]), [len(cht), 2, 0, 1], g.init, 50, g.t_correct, param, initial, eps)


@pytest.mark.parametrize("gridsize", [40, 20, 100])
def test_other_grid_sizes(api, changepoint, gridsize):
    """k = 50 (two lane passes in the trajectory kernel), k = 100 (> 64: per-interval dt from shared memory),
    k = 20."""
    g = changepoint
    rng = np.random.default_rng(3 + gridsize)
    k = 2000 // gridsize
    init = make_init(g, g.times, gridsize)
    cht = g.times[::gridsize][1:-1][:: max(1, k // 10)]
    param, initial, eps = draws(rng, 16, len(cht), k, eps_scale=0.2 * np.sqrt(gridsize / 50))
    x_r, x_i = np.concatenate([[1e6], cht]), [len(cht), 2, 0, 1]
    pb, res = compare(api, g.times, x_r, x_i, init, gridsize, g.t_correct, param, initial, eps)
    # trajectory ESS on this grid
    ok = np.nonzero(res["coalLog"] != orc.SENTINEL)[0][:6]
    nrm, unf = rng.standard_normal((len(ok), k * 2)), rng.random((len(ok), 22))
    d = api.ESlice_general_NC(eps[ok], res["Ode"][ok], {"A": res["FT"]["A"][ok], "Sigma": res["FT"]["Sigma"][ok]},
                              initial[0], init, res["betaN"][ok], g.t_correct, 1.0, res["coalLog"][ok], gridsize, True,
                              "SIR", normals=nrm, uniforms=unf)
    for j, b in enumerate(ok):
        o = orc.ESlice_general_NC(eps[b], res["Ode"][b], {"A": res["FT"]["A"][b], "Sigma": res["FT"]["Sigma"][b]},
                                  initial[0], init, res["betaN"][b], g.t_correct, 1.0, res["coalLog"][b], gridsize, True,
                                  "SIR", normals=nrm[j], uniforms=unf[j])
        assert d["n_evals"][j] == o["n_evals"], (gridsize, j)
        assert d["CoalLog"][j] == pytest.approx(o["CoalLog"], rel=1e-9)
        assert np.max(np.abs(d["OriginTraj"][j] - o["OriginTraj"])) < 1e-11


def test_single_item_and_empty_batch(api, changepoint):
    import ctypes as C
    from lnaphylodyn_b200 import _lib
    g = changepoint
    par = g.final["par"].ravel()
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    r1 = pb.full_loglik(par[2:], par[:2], g.final["OriginTraj"])
    rB = pb.full_loglik(np.tile(par[2:], (3, 1)), np.tile(par[:2], (3, 1)), np.tile(g.final["OriginTraj"], (3, 1, 1)))
    assert r1["coalLog"] == rB["coalLog"][0] == rB["coalLog"][2]
    assert _lib.lib().lna_full_loglik(pb.h, 0, None, None, None, None, None, None, None, None, None, None) == 0
    # errors are loud and descriptive
    pb2 = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, None)
    with pytest.raises(_lib.LnaError, match="without a genealogy"):
        pb2.full_loglik(par[2:], par[:2], g.final["OriginTraj"])
    with pytest.raises(_lib.LnaError, match="multiple of gridsize"):
        api.Problem(g.times[:-1], g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    with pytest.raises(_lib.LnaError, match="hard-wired to SIR"):   # the samplers are p = 2 only, like the reference's
        pbs = api.problem_for(g.times, g.x_r, [39, 3, 0, 2], g.gridsize, g.t_correct, g.init, "SEIR")
        _lib.check(_lib.lib().lna_chains_create(pbs.h, 4) and 0 or -1, "lna_chains_create")


@pytest.mark.parametrize("gridsize,nch", [(8, 39), (40, 0), (50, 3), (1, 2)])
def test_pipelined_full_evaluation_odd_shapes(changepoint, monkeypatch, gridsize, nch):
    """csrc/lna_split.cuh on shapes far from the vignette's: intervals shorter than a ring piece (g = 8, g = 1), more
    intervals than the constant-bank table holds (k = 250 > 64, k = 500), no change points, change points inside a piece,
    ragged batches.  The three-warp pipeline must return the one-warp kernel's bits for every output."""
    from lnaphylodyn_b200 import api
    g = changepoint
    rng = np.random.default_rng(gridsize * 100 + nch)
    nt = 2001 if gridsize > 1 else 501
    times = g.times[:nt]
    cht = np.sort(rng.uniform(times[0], times[-1], nch))
    x_r, x_i = np.concatenate([[1e6], cht]), [nch, 2, 0, 1]
    k = (nt - 1) // gridsize
    pb = api.problem_for(times, x_r, x_i, gridsize, 0.0, None)
    for B in (3, 70):
        par = np.empty((B, 2 + nch + 1))
        par[:, 0] = 2.2 * np.exp(0.1 * rng.standard_normal(B))
        par[:, 1] = 0.2 * np.exp(0.1 * rng.standard_normal(B))
        par[:, 2:2 + nch] = np.exp(0.05 * rng.standard_normal((B, nch)))
        par[:, -1] = 20.0
        ini = np.tile([1e6, 10.0], (B, 1))
        eps = 0.5 * rng.standard_normal((B, k, 2))
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("LNA_SPLIT", mode)
            out[mode] = pb.full_loglik(par, ini, eps, want=("FT", "Ode", "betaN", "LatentTraj"))
        for key in ("Ode", "betaN", "LatentTraj", "status"):
            assert np.array_equal(out["0"][key], out["1"][key]), key
        assert np.array_equal(out["0"]["FT"]["A"], out["1"]["FT"]["A"]) and np.array_equal(out["0"]["FT"]["Sigma"], out["1"]["FT"]["Sigma"])
        assert np.all(np.isfinite(out["1"]["Ode"]))


def test_problem_create_rejects_layouts_the_reference_cannot_run(api, changepoint):
    """x_i[2], x_i[3] are parameter indices AND the S / I columns of the coalescent likelihood: out-of-range values are
    refused at create time, and the fused evaluation only takes the layouts whose rhs and param_transform agree in the
    reference, (nch, 2, 0, 1) for SIR and (nch, 3, 0, 2) for SEIR."""
    from lnaphylodyn_b200._lib import LnaError
    g = changepoint
    for bad in ([39, 2, 0, 2], [39, 2, -1, 1], [39, 2, 3, 1]):
        with pytest.raises(LnaError):
            api.Problem(g.times, g.x_r, bad, g.gridsize, g.t_correct, g.init)
    pb = api.Problem(g.times, g.x_r, [39, 2, 1, 0], g.gridsize, g.t_correct, g.init)   # in range, but not the reference's layout
    par = g.final["par"].ravel()
    with pytest.raises(LnaError, match="must be"):
        pb.full_loglik(par[2:], par[:2], g.final["OriginTraj"])


@pytest.mark.parametrize("gridsize,nch", [(50, 39), (8, 5), (40, 0), (1, 2)])
def test_balanced_full_kernel_returns_the_plain_kernels_bits(changepoint, monkeypatch, gridsize, nch):
    """csrc/lna_kernels.cuh k_full_loglik_bal: the warp-loads of an SM are cut into four equal shares at LNA-interval
    boundaries and a cut load is evaluated in pieces by warps on neighbouring sub-partitions, the running state handed over
    through shared memory.  Small batches (a load cut three times: every warp waits for and hands to a neighbour), ragged
    ones, batches where the two load counts per SM differ, with and without stored outputs, change points inside a piece:
    every output must carry the plain kernel's bits; with the vignette's genealogy the likelihood and its sentinels too."""
    from lnaphylodyn_b200 import api
    g = changepoint
    rng = np.random.default_rng(gridsize * 1000 + nch)
    vignette = (gridsize, nch) == (50, 39)
    nt = 2001 if gridsize > 1 else 501
    times = g.times[:nt]
    if vignette:
        x_r, x_i, init, tc = g.x_r, g.x_i, g.init, g.t_correct
    else:
        cht = np.sort(rng.uniform(times[0], times[-1], nch))
        x_r, x_i, init, tc = np.concatenate([[1e6], cht]), [nch, 2, 0, 1], None, 0.0
    k = (nt - 1) // gridsize
    pb = api.problem_for(times, x_r, x_i, gridsize, tc, init)
    monkeypatch.setenv("LNA_SPLIT", "0")
    for B in (1, 37, 4737, 9489) if vignette else (3, 700):
        par, ini, eps = draws(rng, B, nch, k, eps_scale=0.5)
        if vignette:
            fp = g.final["par"].ravel()
            par[:, 0], par[:, 1], ini = fp[2] * np.exp(0.05 * rng.standard_normal(B)), fp[3], np.tile(fp[:2], (B, 1))
        want = ("coalLog", "FT", "Ode", "betaN", "LatentTraj") if vignette else ("FT", "Ode", "betaN", "LatentTraj")
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("LNA_BAL", mode)
            out[mode] = pb.full_loglik(par, ini, eps, want=want if B < 5000 else want[:1])
        for key in out["0"]:
            if key == "FT":
                assert np.array_equal(out["0"]["FT"]["A"], out["1"]["FT"]["A"]) and np.array_equal(out["0"]["FT"]["Sigma"], out["1"]["FT"]["Sigma"])
            else:
                assert np.array_equal(out["0"][key], out["1"][key]), (key, B)
        if vignette:
            assert np.sum(out["1"]["coalLog"] > orc.SENTINEL) >= B // 4


def test_balanced_full_kernel_against_the_oracle(api, changepoint, monkeypatch):
    """The balanced kernel forced on (LNA_BAL=1, pipeline off) on a batch where every load is cut: oracle parity at 1e-10."""
    g = changepoint
    rng = np.random.default_rng(77)
    monkeypatch.setenv("LNA_SPLIT", "0")
    monkeypatch.setenv("LNA_BAL", "1")
    B = 48
    par = np.tile(g.final["par"].ravel()[2:], (B, 1))
    par[:, 0] *= np.exp(0.05 * rng.standard_normal(B))
    par[:, 2:-1] *= np.exp(0.03 * rng.standard_normal((B, par.shape[1] - 3)))
    ini = np.tile(g.final["par"].ravel()[:2], (B, 1))
    eps = np.tile(g.final["OriginTraj"], (B, 1, 1)) + 0.1 * rng.standard_normal((B, 40, 2))
    compare(api, g.times, g.x_r, g.x_i, g.init, g.gridsize, g.t_correct, par, ini, eps)
