"""The reference-side binding of the LEGACY copies (integration/rcpp_shim_legacy.cpp -> C ABI -> sm_100a kernels), compiled
under the same harness as the reference's own bodies (oracle/ref_harness_legacy.cpp) and run beside them:
    oracle/_ref/liblna_ref.so     harness + src/LNA_NS_general.cpp, src/generalLNA.cpp (class Foo), src/SIRS_period.cpp   (CPU)
    oracle/_ref/liblna_dropin.so  harness + integration/rcpp_shim_legacy.cpp + the CUDA library                          (GPU)
Values within 1e-10 relative; slice-shrink counts and the position the call leaves the session's generator at identical
(for `reps` > 1 the binding draws randn / runif per repetition in the reference's order)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.dropin_available(), reason="oracle/_ref/liblna_dropin.so not built"),
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref/liblna_ref.so not built")]
TOL = 1e-10


@pytest.fixture(scope="module")
def dp():
    return ref.dropin()


def legacy_problem(g, scale=0.0, seed=0):
    rng = np.random.default_rng(seed)
    par = g.final["par"].ravel().copy()
    if scale:
        par[2:4] *= np.exp(scale * rng.standard_normal(2))
        par[4:-1] *= np.exp(0.2 * scale * rng.standard_normal(len(par) - 5))
    return par[2:-1].copy(), g.x_r, [int(g.x_i[0]), 2], par[:2].copy()


def test_lna_ns_general_exports_through_the_binding(dp, changepoint):
    g = changepoint
    param, x_r, x_i, initial = legacy_problem(g, 0.1, 4)
    ts = np.sort(np.concatenate([g.times[::97], x_r[1:4]]))
    assert relerr(dp.betafs(ts, param, x_r, x_i), ref.betafs(ts, param, x_r, x_i)) < 1e-14
    st = np.array([0.9 * x_r[0], 321.0])
    assert relerr(dp.ODE_general_one(st, param, 12.5, x_r, x_i), ref.ODE_general_one(st, param, 12.5, x_r, x_i)) < 1e-14
    assert relerr(dp.general_F(st, param, 12.5, x_r, x_i), ref.general_F(st, param, 12.5, x_r, x_i)) < 1e-14
    a, b = dp.General_ODE_rk45(initial, g.times, param, x_r, x_i), ref.General_ODE_rk45(initial, g.times, param, x_r, x_i)
    assert relerr(a, b) < TOL
    fa, fb = dp.General_IntSigmaF(b[100:150], param, x_r, x_i), ref.General_IntSigmaF(b[100:150], param, x_r, x_i)
    assert relerr(fa["expF"], fb["expF"]) < TOL and np.max(np.abs(fa["Simga"] - fb["Simga"])) < TOL * np.max(np.abs(fb["Simga"]))
    ka, kb = dp.General_KOM_Filter(b, param, g.gridsize, x_r, x_i), ref.General_KOM_Filter(b, param, g.gridsize, x_r, x_i)
    assert relerr(ka["A"], kb["A"]) < TOL and np.max(np.abs(ka["Sigma"] - kb["Sigma"])) < TOL * np.max(np.abs(kb["Sigma"]))
    co = b[:: g.gridsize]
    rng = np.random.default_rng(2)
    foo_d, foo_r = (m.Foo(np.zeros((3, 2)), np.eye(3, 2), [1, 1, 0], [0], [1.0]) for m in (dp, ref))
    for s in range(3):
        z = rng.standard_normal(80)
        for sim_d, sim_r in ((dp.Traj_sim_general, ref.Traj_sim_general), (foo_d.LNA_Traj_sim, foo_r.LNA_Traj_sim)):
            a, c = sim_d(co, kb, g.t_correct, normals=z), sim_r(co, kb, g.t_correct, normals=z)
            assert relerr(a["SimuTraj"][:, 1:], c["SimuTraj"][:, 1:]) < TOL
            assert a["loglike"] == pytest.approx(c["loglike"], rel=1e-9)
        assert dp.log_like_traj_general(c["SimuTraj"], co, kb, 50, g.t_correct) == pytest.approx(
            ref.log_like_traj_general(c["SimuTraj"], co, kb, 50, g.t_correct), rel=1e-9)
        assert foo_d.LNA_log_like_traj(c["SimuTraj"], co, kb, 50, g.t_correct) == pytest.approx(
            foo_r.LNA_log_like_traj(c["SimuTraj"], co, kb, 50, g.t_correct), rel=1e-9)
        a = dp.Traj_sim_ezG(initial, g.times, param, 50, x_r, x_i, g.t_correct, normals=z)
        c = ref.Traj_sim_ezG(initial, g.times, param, 50, x_r, x_i, g.t_correct, normals=z)
        assert relerr(a["SimuTraj"][:, 1:], c["SimuTraj"][:, 1:]) < 1e-9


@pytest.mark.parametrize("volz", [True, False])
def test_eslice_general_and_foo_through_the_binding(dp, changepoint, volz):
    g = changepoint
    param, x_r, x_i, initial = legacy_problem(g, 0.05, 1)
    ode = orc.General_ODE_rk45(initial, g.times, param, x_r, x_i)
    kf = orc.General_KOM_Filter(ode, param, g.gridsize, x_r, x_i)
    co = ode[:: g.gridsize]
    betaN = orc.betafs(co[:, 0], param, x_r, x_i)
    rng = np.random.default_rng(8)
    f_cur = orc.Traj_sim_general(co, kf, g.t_correct, normals=0.3 * rng.standard_normal(80))["SimuTraj"]
    lg = f_cur.copy()
    lg[:, 1:] = np.log(lg[:, 1:])
    assert dp.volz_loglik_nh(g.init, lg, betaN, g.t_correct) == pytest.approx(ref.volz_loglik_nh(g.init, lg, betaN, g.t_correct), rel=1e-12)
    foo_d, foo_r = (m.Foo(np.zeros((3, 2)), np.eye(3, 2), [1, 1, 0], [0], [1.0]) for m in (dp, ref))
    total = 0
    for reps in (1, 3):
        for s in range(5):
            kw = dict(normals=rng.standard_normal(80 * reps), uniforms=rng.uniform(size=22 * reps))
            args = (f_cur, co, kf, initial, g.init, betaN, g.t_correct, 25.0, reps, 50, volz)
            for fd, fr in ((dp.ESlice_general, ref.ESlice_general), (foo_d.LNA_ESlice, foo_r.LNA_ESlice)):
                a, b = fd(*args, **kw), fr(*args, **kw)
                assert a["n_uniforms"] == b["n_uniforms"]   # the generator is left where the reference leaves it
                assert relerr(a["newTraj"][:, 1:], b["newTraj"][:, 1:]) < TOL
            total += b["n_uniforms"]
    assert total > 2 * 20
    u, z = rng.uniform(size=22), 1e13 * rng.standard_normal(80)
    a = dp.ESlice_general(f_cur, co, kf, initial, g.init, betaN, g.t_correct, 25.0, 1, 50, volz, normals=z, uniforms=u)
    assert a["n_uniforms"] == 22 and np.array_equal(a["newTraj"], f_cur)


def test_sirs_samplers_through_the_binding(dp, changepoint):
    t = np.linspace(0, 100, 2001)
    param, period = [3e-7, 0.2, 0.05, 0.3], 40.0
    initial = np.array([1e6, 10.0, 0.0])
    ode = orc.ODE2(initial, t, param, period)
    kf = orc.SIRS_KOM_Filter(ode, param, 50, period)
    kf = {"A": kf["A"], "Sigma": kf["Sigma"] + 1e-6 * np.eye(3)[:, :, None]}
    co = ode[::50]
    g = changepoint
    rng = np.random.default_rng(5)
    params4 = [param[0], param[3], period, 1.0]
    z = 0.2 * rng.standard_normal(120)
    a, b = dp.Traj_sim_SIRS(co[0, 1:], co, kf, 90.0, normals=z), ref.Traj_sim_SIRS(co[0, 1:], co, kf, 90.0, normals=z)
    assert relerr(a["SimuTraj"][:, 1:], b["SimuTraj"][:, 1:]) < 1e-9
    assert a["loglike"] == pytest.approx(b["loglike"], rel=1e-6)
    f_cur = b["SimuTraj"].copy()
    f_cur[0, 0] = co[0, 0]
    for reps in (1, 2):
        for s in range(4):
            kw = dict(normals=rng.standard_normal(120 * reps), uniforms=rng.uniform(size=22 * reps))
            args = (f_cur, co, kf, co[0, 1:], g.init, params4, 90.0, 10.0, reps, 50, True)
            a, b = dp.ESlice_SIRS(*args, **kw), ref.ESlice_SIRS(*args, **kw)
            assert a["n_uniforms"] == b["n_uniforms"]
            assert relerr(a["newTraj"][:, 1:], b["newTraj"][:, 1:]) < 1e-9


def test_foo_through_the_binding(dp):
    A_pre = np.array([[1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    A_post = np.array([[0.0, 2.0], [0.0, 0.0], [1.0, 0.0]])
    param, N = [2.2, 0.2, 1e-4], 1e5
    fd, fr = dp.Foo(A_pre, A_post, param, [0], [N]), ref.Foo(A_pre, A_post, param, [0], [N])
    st = np.array([9e4, 250.0])
    for m in ("rate_h", "dh", "ODE_ones"):
        assert np.allclose(getattr(fd, m)(st), getattr(fr, m)(st), rtol=1e-14, atol=0)
    t = np.linspace(0, 100, 2001)
    assert relerr(fd.ODE_rk45([N - 1, 1.0], t), fr.ODE_rk45([N - 1, 1.0], t)) < TOL
    # Foo::IntSigmaF throws in both builds (F_vec's arma::vec return type; the empty local `A`)
    import ctypes as C
    for lib in (ref._get_lib()._dll, dp.lib()._dll):
        rc = lib.ref_foo_intsigmaf_runs(orc._p(orc._f(A_pre)), orc._p(orc._f(A_post)), orc._p(orc._d(param)), C.c_double(N),
                                        orc._p(orc._f(np.zeros((50, 3)))), 50)
        assert rc != 0
