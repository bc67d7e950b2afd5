"""The oracle's restatement of the reference's LEGACY copies of the pipeline against the reference code as
written (CPU only): src/LNA_NS_general.cpp, class Foo of src/generalLNA.cpp and ESlice_SIRS of src/SIRS_period.cpp, compiled
verbatim into oracle/_ref/liblna_ref.so (Foo's .cpp is #included by oracle/ref_harness_legacy.cpp; RCPP_MODULE is a no-op).
Tolerance 1e-12 relative on values; slice decisions (uniforms consumed, evaluations, cap) identical."""
import numpy as np
import pytest

from conftest import Golden, relerr
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/liblna_ref.so not built and /root/reference absent")


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    assert float(np.max(np.abs(a - b))) <= tol * scale


def legacy_problem(g, scale=0.0, seed=0):
    """(param, x_r, x_i2, initial) of the legacy layout: param = (R0, gamma, ch_1..ch_nch), x_i = (nch, 2)."""
    rng = np.random.default_rng(seed)
    par = g.final["par"].ravel().copy()
    if scale:
        par[2:4] *= np.exp(scale * rng.standard_normal(2))
        par[4:-1] *= np.exp(0.2 * scale * rng.standard_normal(len(par) - 5))
    return par[2:-1].copy(), g.x_r, [int(g.x_i[0]), 2], par[:2].copy()


@pytest.mark.parametrize("name", ["changepoint_sim", "constant_sim"])
def test_lna_ns_general_integrator_and_moments(name):
    g = Golden(name)
    for seed in range(3):
        param, x_r, x_i, initial = legacy_problem(g, 0.1 * seed, seed)
        ts = np.sort(np.concatenate([g.times[::97], x_r[1:4], x_r[1:4] - 1e-9]))
        close(ref.betafs(ts, param, x_r, x_i), orc.betafs(ts, param, x_r, x_i), 1e-14)
        assert ref.betaf(ts[3], param, x_r, x_i) == orc.betaf(ts[3], param, x_r, x_i)
        st = np.array([0.9 * x_r[0], 321.0])
        close(ref.ODE_general_one(st, param, 12.5, x_r, x_i), orc.ODE_general_one(st, param, 12.5, x_r, x_i), 1e-15)
        close(ref.general_F(st, param, 12.5, x_r, x_i), orc.general_F(st, param, 12.5, x_r, x_i), 1e-15)
        a, b = ref.General_ODE_rk45(initial, g.times, param, x_r, x_i), orc.General_ODE_rk45(initial, g.times, param, x_r, x_i)
        close(a, b, 1e-13)
        # the legacy integrator IS the current one on the SIR model (same stages, same change-point scan)
        close(b, orc.ODE_rk45(initial, g.times, np.append(param, 1.0), x_r, g.x_i), 0.0)
        fa, fb = ref.General_IntSigmaF(b[100:150], param, x_r, x_i), orc.General_IntSigmaF(b[100:150], param, x_r, x_i)
        close(fa["expF"], fb["expF"])
        close(fa["Simga"], fb["Simga"])
        ka, kb = ref.General_KOM_Filter(b, param, g.gridsize, x_r, x_i), orc.General_KOM_Filter(b, param, g.gridsize, x_r, x_i)
        close(ka["A"], kb["A"])
        close(ka["Sigma"], kb["Sigma"])


def _centred_setup(g, seed=0):
    param, x_r, x_i, initial = legacy_problem(g, 0.05, seed)
    ode = orc.General_ODE_rk45(initial, g.times, param, x_r, x_i)
    kf = orc.General_KOM_Filter(ode, param, g.gridsize, x_r, x_i)
    co = ode[:: g.gridsize]
    return param, x_r, x_i, initial, ode, kf, co, orc.betafs(co[:, 0], param, x_r, x_i)


def test_legacy_centred_simulation_and_density(changepoint):
    """Traj_sim_general / log_like_traj_general / Traj_sim_ezG and their Foo twins (inv2 density, chols noise)."""
    g = changepoint
    param, x_r, x_i, initial, ode, kf, co, betaN = _centred_setup(g)
    rng = np.random.default_rng(3)
    foo_r, foo_o = ref.Foo(np.zeros((3, 2)), np.eye(3, 2), [1, 1, 0], [0], [1.0]), orc.Foo(np.zeros((3, 2)), np.eye(3, 2), [1, 1, 0], [0], [1.0])
    for s in range(5):
        z = rng.standard_normal(80)
        a, b = ref.Traj_sim_general(co, kf, g.t_correct, normals=z), orc.Traj_sim_general(co, kf, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], b["SimuTraj"][:, 1:])
        assert a["loglike"] == pytest.approx(b["loglike"], rel=1e-10)
        c = foo_r.LNA_Traj_sim(co, kf, g.t_correct, normals=z)
        close(c["SimuTraj"][:, 1:], b["SimuTraj"][:, 1:])
        assert c["loglike"] == pytest.approx(b["loglike"], rel=1e-10)
        la = ref.log_like_traj_general(b["SimuTraj"], co, kf, 50, g.t_correct)
        lb = orc.log_like_traj_general(b["SimuTraj"], co, kf, 50, g.t_correct)
        lc = foo_r.LNA_log_like_traj(b["SimuTraj"], co, kf, 50, g.t_correct)
        assert la == pytest.approx(lb, rel=1e-10) and lc == pytest.approx(lb, rel=1e-10)
        assert foo_o.LNA_log_like_traj(b["SimuTraj"], co, kf, 50, g.t_correct) == lb
        # the simulator's own density and the density of its output agree (both go through inv2)
        assert lb == pytest.approx(b["loglike"], rel=1e-6)
        a = ref.Traj_sim_ezG(initial, g.times, param, 50, x_r, x_i, g.t_correct, normals=z)
        c = orc.Traj_sim_ezG(initial, g.times, param, 50, x_r, x_i, g.t_correct, normals=z)
        close(a["SimuTraj"][:, 1:], c["SimuTraj"][:, 1:])
        assert a["loglike"] == pytest.approx(c["loglike"], rel=1e-10)


@pytest.mark.parametrize("volz", [True, False])
def test_eslice_general_and_foo_step_for_step(changepoint, volz):
    """ESlice_general (LNA_NS_general.cpp:320-394) and Foo::LNA_ESlice, reps = 1 and 3, both likelihoods."""
    g = changepoint
    param, x_r, x_i, initial, ode, kf, co, betaN = _centred_setup(g, 1)
    rng = np.random.default_rng(8)
    foo_r = ref.Foo(np.zeros((3, 2)), np.eye(3, 2), [1, 1, 0], [0], [1.0])
    f_cur = orc.Traj_sim_general(co, kf, g.t_correct, normals=0.3 * rng.standard_normal(80))["SimuTraj"]
    assert np.min(f_cur[:, 1:]) >= 1.0
    # volz_loglik_nh on the log trajectory
    lg = f_cur.copy()
    lg[:, 1:] = np.log(lg[:, 1:])
    assert ref.volz_loglik_nh(g.init, lg, betaN, g.t_correct) == pytest.approx(orc.volz_loglik_nh(g.init, lg, betaN, g.t_correct), rel=1e-12)
    evals = []
    for reps in (1, 3):
        for s in range(6):
            kw = dict(normals=rng.standard_normal(80 * reps), uniforms=rng.uniform(size=22 * reps))
            args = (f_cur, co, kf, initial, g.init, betaN, g.t_correct, 25.0, reps, 50, volz)
            a, b, c = ref.ESlice_general(*args, **kw), orc.ESlice_general(*args, **kw), foo_r.LNA_ESlice(*args, **kw)
            assert a["n_uniforms"] == b["n_uniforms"] == c["n_uniforms"]
            close(a["newTraj"], b["newTraj"])
            close(c["newTraj"], b["newTraj"])
            evals.append(b["n_evals"])
    assert max(evals) > 2 * 3  # some shrinking happened
    # the 20-shrink cap: a proposal so wild that every rotated trajectory has a negative state -> back to f_cur, bit for bit
    u = rng.uniform(size=22)
    z = 1e13 * rng.standard_normal(80)
    a, b = ref.ESlice_general(f_cur, co, kf, initial, g.init, betaN, g.t_correct, 25.0, 1, 50, volz, normals=z, uniforms=u), \
        orc.ESlice_general(f_cur, co, kf, initial, g.init, betaN, g.t_correct, 25.0, 1, 50, volz, normals=z, uniforms=u)
    assert b["status"] & 8 and b["n_uniforms"] == 22 == a["n_uniforms"]
    assert np.array_equal(b["newTraj"], f_cur) and np.array_equal(a["newTraj"], f_cur)


def test_eslice_sirs_step_for_step():
    """ESlice_SIRS (SIRS_period.cpp:313-402): p = 3, proposal Traj_sim_SIRS(state, .), betat = betaDyn(.) * N; wherever the
    jittered Cholesky of the rank-2 covariance exists for the oracle it must for the reference, and the chains agree."""
    t = np.linspace(0, 100, 2001)
    param, period = [3e-7, 0.2, 0.05, 0.3], 40.0
    initial = np.array([1e6, 10.0, 0.0])
    ode = orc.ODE2(initial, t, param, period)
    kf = orc.SIRS_KOM_Filter(ode, param, 50, period)
    kf = {"A": kf["A"], "Sigma": kf["Sigma"] + 1e-6 * np.eye(3)[:, :, None]}  # lifted off the singular plane: chol() exists
    co = ode[::50]
    g = Golden("changepoint_sim")
    rng = np.random.default_rng(5)
    params4 = [param[0], param[3], period, 1.0]
    f_cur = orc.Traj_sim_SIRS(co[0, 1:], co, kf, 90.0, normals=0.2 * rng.standard_normal(120))["SimuTraj"]
    f_cur[0, 0] = co[0, 0]
    assert np.min(f_cur[:, 1:3]) > 1.0
    n = 0
    for reps in (1, 2):
        for s in range(5):
            kw = dict(normals=rng.standard_normal(120 * reps), uniforms=rng.uniform(size=22 * reps))
            args = (f_cur, co, kf, co[0, 1:], g.init, params4, 90.0, 10.0, reps, 50, True)
            a, b = ref.ESlice_SIRS(*args, **kw), orc.ESlice_SIRS(*args, **kw)
            assert a["n_uniforms"] == b["n_uniforms"]
            close(a["newTraj"], b["newTraj"], 1e-11)
            n += b["n_evals"]
    assert n > 20


def test_foo_reaction_network():
    """class Foo: rate_h, dh, ODE_ones, ODE_rk45 (SIR with an immigration reaction); IntSigmaF cannot run as written."""
    A_pre = np.array([[1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    A_post = np.array([[0.0, 2.0], [0.0, 0.0], [1.0, 0.0]])
    param, N = [2.2, 0.2, 1e-4], 1e5
    fr, fo = ref.Foo(A_pre, A_post, param, [0], [N]), orc.Foo(A_pre, A_post, param, [0], [N])
    st = np.array([9e4, 250.0])
    for m in ("rate_h", "dh", "ODE_ones"):
        close(getattr(fr, m)(st), getattr(fo, m)(st), 1e-15)
    t = np.linspace(0, 100, 2001)
    a, b = fr.ODE_rk45([N - 1, 1.0], t), fo.ODE_rk45([N - 1, 1.0], t)
    close(a, b, 1e-13)
    assert b[-1, 2] > 1.0 and abs(b[:, 1:].sum(1)[-1] - N) > 1.0  # (immigration: the population grows)
    # Foo::F_vec returns a 2 x 2 product through an arma::vec and IntSigmaF multiplies by an EMPTY local `A` (:130, 146):
    # neither can run as written; the oracle offers the product F_vec computes under its own name
    close(fo.F_matrix(st), (A_post - A_pre).T @ fo.dh(st), 1e-15)
    lib = ref._get_lib()._dll
    import ctypes as C
    rc = lib.ref_foo_intsigmaf_runs(orc._p(orc._f(A_pre)), orc._p(orc._f(A_post)), orc._p(orc._d(param)), C.c_double(N),
                                    orc._p(orc._f(b[:50])), 50)
    assert rc != 0, "Foo::IntSigmaF ran on the stand-in headers although Armadillo would reject it"
