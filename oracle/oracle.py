"""ctypes binding of the CPU oracle (oracle/liblna_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/lna_oracle.h.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module.  The function names
mirror the reference's Rcpp exports (R/RcppExports.R) so that tests read like the reference's.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblna_oracle.so")

MODELS = {"SIR": 0, "SEIR": 1, "SEIR2": 2}
MODEL_P = {"SIR": 2, "SEIR": 3, "SEIR2": 2}
TRANSX = {"standard": 0, "log": 1}
SENTINEL = -10000000.0

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Cfg(C.Structure):
    _fields_ = [("model", C.c_int), ("transX", C.c_int), ("p", C.c_int), ("nparam_total", C.c_int),
                ("x_r", _dp), ("x_i", C.c_int * 4)]


class _Init(C.Structure):
    _fields_ = [("E", C.c_int), ("ng", C.c_int), ("C", _dp), ("D", _dp), ("y", _dp), ("gridrep", _dp)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("lna_oracle.c", "lna_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liblna_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        for name in ("orc_volz_loglik_nh2", "orc_volz_loglik_nh3", "orc_coal_loglik3", "orc_coal_loglik",
                     "orc_log_like_traj_general2", "orc_ode_rk45_stop", "orc_volz_loglik_nh"):
            getattr(_lib, name).restype = C.c_double
    return _lib


def _d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _f(a):
    """column-major double copy"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def _ints(a):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).round().astype(np.int32))
    return arr, arr.ctypes.data_as(_ip)


class Cfg:
    """(model, transX, x_r, x_i) pack; keeps the numpy buffers alive."""

    def __init__(self, x_r, x_i, nparam_total, model="SIR", transX="standard"):
        self.x_r = _d(x_r)
        self.x_i = [int(round(float(v))) for v in np.asarray(x_i).ravel()]
        self.model, self.transX = model, transX
        self.p = MODEL_P[model]
        self.nparam_total = int(nparam_total)
        self.c = _Cfg(MODELS[model], TRANSX[transX], self.p, self.nparam_total, _p(self.x_r),
                      (C.c_int * 4)(*self.x_i))

    @property
    def ref(self):
        return C.byref(self.c)


class Init:
    """Positional view of coal_lik_init()'s list: needs C, D, y, gridrep, ng."""

    def __init__(self, init):
        g = (lambda k: init[k]) if not hasattr(init, "files") else (lambda k: init[k])
        self.C, self.D, self.y = _d(g("C")), _d(g("D")), _d(g("y"))
        self.gridrep = _d(g("gridrep"))
        self.ng = int(np.asarray(g("ng")).ravel()[0])
        self.E = len(self.C)
        self.c = _Init(self.E, self.ng, _p(self.C), _p(self.D), _p(self.y), _p(self.gridrep))

    @property
    def ref(self):
        return C.byref(self.c)


# ----------------------------------------------------------------------------- Rcpp-named wrappers

def betaTs(param, times, x_r, x_i):
    param, times, x_r = _d(param), _d(times), _d(x_r)
    xi, xip = _ints(x_i)
    out = np.empty(len(times))
    lib().orc_betaTs(_p(param), _p(times), len(times), _p(x_r), xip, _p(out))
    return out


def param_transform(t, param, x_r, x_i):
    param, x_r = _d(param), _d(x_r)
    xi, xip = _ints(x_i)
    out = np.empty(len(param))
    lib().orc_param_transform(C.c_double(t), _p(param), len(param), _p(x_r), xip, _p(out))
    return out


def ODE_rk45(initial, t, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    initial, t, param = _d(initial), _d(t), _d(param)
    out = np.empty((len(t), cfg.p + 1), order="F")
    lib().orc_ode_rk45(cfg.ref, _p(initial), _p(t), len(t), _p(param), _p(out))
    return out


def ODE_rk45_stop(initial, t, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard", tol=5):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    initial, t, param = _d(initial), _d(t), _d(param)
    return lib().orc_ode_rk45_stop(cfg.ref, _p(initial), _p(t), len(t), _p(param), C.c_double(tol))


def SigmaF(Traj_par, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    T, param = _f(Traj_par), _d(param)
    p = cfg.p
    eF, S = np.empty((p, p), order="F"), np.empty((p, p), order="F")
    lib().orc_sigmaF(cfg.ref, _p(T), T.shape[0], T.shape[0], _p(param), _p(eF), _p(S))
    return {"expF": eF, "Sigma": S}


def _cubes(p, k):
    return np.empty((p, p, k), order="F"), np.empty((p, p, k), order="F")


def KF_param(OdeTraj, param, gridsize, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    O, param = _f(OdeTraj), _d(param)
    k = (O.shape[0] - 1) // gridsize
    A, S = _cubes(cfg.p, k)
    lib().orc_kf_param(cfg.ref, _p(O), O.shape[0], _p(param), int(gridsize), _p(A), _p(S))
    return {"A": A, "Sigma": S}


def KF_param_chol(OdeTraj, param, gridsize, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    O, param = _f(OdeTraj), _d(param)
    k = (O.shape[0] - 1) // gridsize
    A, L = _cubes(cfg.p, k)
    st = lib().orc_kf_param_chol(cfg.ref, _p(O), O.shape[0], _p(param), int(gridsize), _p(A), _p(L))
    if st:
        raise ValueError("Invalid input for cholesky decomposition., parametrization fails")
    return {"A": A, "Sigma": L}


def Ode_Coarse_Slicer(Ode_thin, gridsize):
    O = _f(Ode_thin)
    d = O.shape[0] // gridsize + 1
    out = np.empty((d, O.shape[1]), order="F")
    lib().orc_ode_coarse_slicer(_p(O), O.shape[0], O.shape[1], int(gridsize), _p(out))
    return out


def New_Param_List(param, initial, gridsize, t, x_r, x_i, transP="changepoint", model="SIR",
                   transX="standard"):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    param, initial, t = _d(param), _d(initial), _d(t)
    nt, p = len(t), cfg.p
    k, nrow = (nt - 1) // gridsize, nt // gridsize + 1
    A, L = _cubes(p, k)
    ode, betaN = np.empty((nrow, p + 1), order="F"), np.empty(nrow)
    st = lib().orc_new_param_list(cfg.ref, _p(param), _p(initial), int(gridsize), _p(t), nt, _p(A), _p(L),
                                  _p(ode), _p(betaN))
    if st:
        raise ValueError("Invalid input for cholesky decomposition., parametrization fails")
    return {"FT": {"A": A, "Sigma": L}, "Ode": ode, "betaN": betaN}


def TransformTraj(OdeTraj, OriginLatent, Filter_NC):
    O, eps = _f(OdeTraj), _f(OriginLatent)
    A, L = _f(Filter_NC["A"]), _f(Filter_NC["Sigma"])
    out = np.empty_like(O, order="F")
    lib().orc_transform_traj(_p(O), O.shape[0], O.shape[1] - 1, _p(eps), eps.shape[0], _p(A), _p(L), _p(out))
    return out


def Traj_sim_general_noncentral(OdeTraj, Filter_NC, t_correct, normals):
    """`normals`: k*p pre-drawn N(0,1), step-major (step i uses normals[i*p:(i+1)*p])."""
    O = _f(OdeTraj)
    A, L = _f(Filter_NC["A"]), _f(Filter_NC["Sigma"])
    nrow, p = O.shape[0], O.shape[1] - 1
    nrm = _d(normals)
    traj, origin = np.empty_like(O, order="F"), np.empty((nrow - 1, p), order="F")
    lm, lo = C.c_double(), C.c_double()
    lib().orc_traj_sim_noncentral(_p(O), nrow, p, _p(A), _p(L), C.c_double(t_correct), _p(nrm), _p(traj),
                                  _p(origin), C.byref(lm), C.byref(lo))
    return {"SimuTraj": traj, "OriginTraj": origin, "logMultiNorm": lm.value, "logOrigin": lo.value}


def log_like_traj_general2(SdeTraj, OdeTraj, Filter, gridsize, t_correct):
    S, O = _f(SdeTraj), _f(OdeTraj)
    A, Sg = _f(Filter["A"]), _f(Filter["Sigma"])
    return lib().orc_log_like_traj_general2(_p(S), _p(O), S.shape[0], S.shape[1] - 1, _p(A), _p(Sg),
                                            C.c_double(t_correct))


def volz_loglik_nh2(init, f1, betaN, t_correct, index, transX="standard", _fn="orc_volz_loglik_nh2"):
    ini = init if isinstance(init, Init) else Init(init)
    F, b = _f(f1), _d(betaN)
    ix, ixp = _ints(index)
    return getattr(lib(), _fn)(ini.ref, _p(F), F.shape[0], F.shape[1] - 1, _p(b), C.c_double(t_correct), ixp,
                               TRANSX[transX])


def volz_loglik_nh3(init, f1, betaN, t_correct, index, transX="standard"):
    return volz_loglik_nh2(init, f1, betaN, t_correct, index, transX, _fn="orc_volz_loglik_nh3")


def coal_loglik3(init, f1, t_correct, lam, Index, transX="standard"):
    ini = init if isinstance(init, Init) else Init(init)
    F = _f(f1)
    return lib().orc_coal_loglik3(ini.ref, _p(F), F.shape[0], F.shape[1] - 1, C.c_double(t_correct),
                                  C.c_double(lam), int(Index), TRANSX[transX])


def coal_loglik(init, f1, t_correct, lam, gridsize=1, transX="standard"):
    ini = init if isinstance(init, Init) else Init(init)
    F = _f(f1)
    return lib().orc_coal_loglik(ini.ref, _p(F), F.shape[0], F.shape[1] - 1, C.c_double(t_correct),
                                 C.c_double(lam))


def full_loglik(cfg, ini, param, initial, t, gridsize, OriginTraj, t_correct):
    """One FULL evaluation (New_Param_List -> TransformTraj -> volz_loglik_nh2). Returns dict + status."""
    param, initial, t, eps = _d(param), _d(initial), _d(t), _f(OriginTraj)
    nt, p = len(t), cfg.p
    k, nrow = (nt - 1) // gridsize, nt // gridsize + 1
    A, L = _cubes(p, k)
    ode, betaN, lat = np.empty((nrow, p + 1), order="F"), np.empty(nrow), np.empty((nrow, p + 1), order="F")
    cl = C.c_double()
    st = lib().orc_full_loglik(cfg.ref, ini.ref, _p(param), _p(initial), _p(t), nt, int(gridsize), _p(eps),
                               C.c_double(t_correct), _p(A), _p(L), _p(ode), _p(betaN), _p(lat), C.byref(cl))
    return {"status": st, "FT": {"A": A, "Sigma": L}, "Ode": ode, "betaN": betaN, "LatentTraj": lat,
            "coalLog": cl.value}


def full_loglik_batch(cfg, ini, param, initial, t, gridsize, OriginTraj, t_correct):
    """B FULL evaluations on ONE core; param (B,npt), initial (B,p), OriginTraj (B,k,p) natural indexing.
    ctypes releases the GIL during the call, so several of these can run on separate threads."""
    param, initial, t = _d(param), _d(initial), _d(t)
    B = param.shape[0]
    eps = np.ascontiguousarray(np.transpose(np.asarray(OriginTraj, dtype=np.float64), (0, 2, 1)))
    cl, st = np.empty(B), np.zeros(B, dtype=np.int32)
    lib().orc_full_loglik_batch(cfg.ref, ini.ref, C.c_long(B), _p(param), _p(initial), _p(t), len(t), int(gridsize),
                                _p(eps), C.c_double(t_correct), _p(cl), st.ctypes.data_as(_ip))
    return cl, st


def Update_Param(param, initial, t, OriginTraj, x_r, x_i, init, gridsize, coal_log=0.0,
                 prior_proposal_offset=0.0, t_correct=0.0, transP="changepoint", model="SIR",
                 transX="standard", volz=True, *, unif):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    ini = init if isinstance(init, Init) else Init(init)
    param, initial, t, eps = _d(param), _d(initial), _d(t), _f(OriginTraj)
    nt, p = len(t), cfg.p
    k, nrow = (nt - 1) // gridsize, nt // gridsize + 1
    A, L = _cubes(p, k)
    ode, betaN, lat = np.empty((nrow, p + 1), order="F"), np.empty(nrow), np.empty((nrow, p + 1), order="F")
    cl, acc = C.c_double(), C.c_int()
    st = lib().orc_update_param(cfg.ref, ini.ref, _p(param), _p(initial), _p(t), nt, _p(eps), int(gridsize),
                                C.c_double(coal_log), C.c_double(prior_proposal_offset), C.c_double(t_correct),
                                C.c_double(unif), C.byref(acc), _p(A), _p(L), _p(ode), _p(betaN), _p(lat),
                                C.byref(cl))
    if st & 1:
        raise ValueError("Invalid input for cholesky decomposition., parametrization fails")
    if not acc.value:
        return {"accept": False, "coalLog_proposed": cl.value}
    return {"accept": True, "FT": {"A": A, "Sigma": L}, "Ode": ode, "betaN": betaN, "coalLog": cl.value,
            "LatentTraj": lat}


def Param_Slice_update(param, x_r, x_i, theta, newChs, rho=1):
    param, newChs = _d(param), _d(newChs)
    xi, xip = _ints(x_i)
    out = np.empty(len(param))
    lib().orc_param_slice_update(_p(param), len(param), xip, C.c_double(theta), _p(newChs), _p(out))
    return out


def _pr8(priorList):
    return _d(np.concatenate([np.asarray(v, dtype=np.float64).ravel()[:2] for v in priorList]))


def Param_Slice_update_all2(par_old, x_r, x_i, theta, newChs, ESS_vec, priorList):
    par_old, newChs = _d(par_old), _d(newChs)
    xi, xip = _ints(x_i)
    ev, evp = _ints(ESS_vec)
    pr = _pr8(priorList)
    out = np.empty(len(par_old))
    lib().orc_param_slice_update_all2(_p(par_old), len(par_old), xip, C.c_double(theta), _p(newChs), evp,
                                      _p(pr), _p(out))
    return out


def ESlice_par_General(par_old, t, OriginTraj, priorList, x_r, x_i, init, gridsize, ESS_vec, coal_log=0.0,
                       t_correct=0.0, transP="changepoint", model="SIR", transX="standard", volz=True, *,
                       normals, uniforms):
    par_old, t, eps = _d(par_old), _d(t), _f(OriginTraj)
    cfg = Cfg(x_r, x_i, len(par_old) - 2, model, transX)
    ini = init if isinstance(init, Init) else Init(init)
    ev, evp = _ints(ESS_vec)
    pr, nrm, unf = _pr8(priorList), _d(normals), _d(uniforms)
    nt, p = len(t), 2
    k, nrow = (nt - 1) // gridsize, nt // gridsize + 1
    A, L = _cubes(p, k)
    ode, betaN, lat = np.empty((nrow, p + 1), order="F"), np.empty(nrow), np.empty((nrow, p + 1), order="F")
    par_new, eps_new = np.empty(len(par_old)), np.empty((k, p), order="F")
    cl, lo = C.c_double(), C.c_double()
    nn, nu, ne = C.c_int(), C.c_int(), C.c_int()
    st = lib().orc_eslice_par_general(cfg.ref, ini.ref, _p(par_old), len(par_old), _p(t), nt, _p(eps), _p(pr),
                                      int(gridsize), evp, C.c_double(coal_log), C.c_double(t_correct),
                                      _p(nrm), _p(unf), C.byref(nn), C.byref(nu), C.byref(ne), _p(par_new),
                                      _p(A), _p(L), _p(ode), _p(betaN), _p(lat), _p(eps_new), C.byref(cl),
                                      C.byref(lo))
    return {"betaN": betaN, "FT": {"A": A, "Sigma": L}, "OdeTraj": ode, "par": par_new, "LatentTraj": lat,
            "CoalLog": cl.value, "OriginTraj": eps_new, "logOrigin": lo.value,
            "status": st, "n_normals": nn.value, "n_uniforms": nu.value, "n_evals": ne.value}


def ESlice_general_NC(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, coal_log=-99999999.0,
                      gridsize=100, volz=False, model="SIR", transX="standard", *, normals, uniforms):
    f, O = _f(f_cur), _f(OdeTraj)
    A, L = _f(FTs["A"]), _f(FTs["Sigma"])
    ini = init if isinstance(init, Init) else Init(init)
    b, nrm, unf = _d(betaN), _d(normals), _d(uniforms)
    k, p = f.shape
    Index = (C.c_int * 2)(0, 1 if model == "SIR" else 2)
    lat, eps_new = np.empty((k + 1, p + 1), order="F"), np.empty((k, p), order="F")
    cl, lo = C.c_double(), C.c_double()
    nn, nu, ne = C.c_int(), C.c_int(), C.c_int()
    st = lib().orc_eslice_general_nc(ini.ref, _p(f), k, p, _p(O), _p(A), _p(L), _p(b), C.c_double(t_correct),
                                     C.c_double(lam), C.c_double(coal_log), int(bool(volz)), Index,
                                     TRANSX[transX], _p(nrm), _p(unf), C.byref(nn), C.byref(nu), C.byref(ne),
                                     _p(lat), _p(eps_new), C.byref(lo), C.byref(cl))
    if st < 0:
        raise NotImplementedError("ESlice_general_NC(volz=False, coal_log=-99999999): the reference hands the increment "
                                  "matrix to coal_loglik3 as a trajectory (:1720); not restated")
    return {"LatentTraj": lat, "OriginTraj": eps_new, "logOrigin": lo.value, "CoalLog": cl.value,
            "status": st, "n_normals": nn.value, "n_uniforms": nu.value, "n_evals": ne.value}


def ESlice_change_points(param, initial, t, OriginTraj, x_r, x_i, init, gridsize, coal_log=0.0, t_correct=0.0,
                         transP="changepoint", model="SIR", transX="standard", volz=True, *, normals, uniforms):
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    ini = init if isinstance(init, Init) else Init(init)
    param, initial, t, eps = _d(param), _d(initial), _d(t), _f(OriginTraj)
    nrm, unf = _d(normals), _d(uniforms)
    nt, p = len(t), cfg.p
    k, nrow = (nt - 1) // gridsize, nt // gridsize + 1
    A, L = _cubes(p, k)
    ode, betaN, lat = np.empty((nrow, p + 1), order="F"), np.empty(nrow), np.empty((nrow, p + 1), order="F")
    param_new = np.empty(len(param))
    cl, lcp = C.c_double(), C.c_double()
    nn, nu, ne = C.c_int(), C.c_int(), C.c_int()
    st = lib().orc_eslice_change_points(cfg.ref, ini.ref, _p(param), _p(initial), _p(t), nt, _p(eps),
                                        int(gridsize), C.c_double(coal_log), C.c_double(t_correct), _p(nrm),
                                        _p(unf), C.byref(nn), C.byref(nu), C.byref(ne), _p(param_new), _p(A),
                                        _p(L), _p(ode), _p(betaN), _p(lat), C.byref(cl), C.byref(lcp))
    return {"betaN": betaN, "FT": {"A": A, "Sigma": L}, "OdeTraj": ode, "param": param_new, "LatentTraj": lat,
            "CoalLog": cl.value, "LogChProb": lcp.value,
            "status": st, "n_normals": nn.value, "n_uniforms": nu.value, "n_evals": ne.value}


def chols(S):
    S = _f(S)
    out = np.empty((2, 2), order="F")
    lib().orc_chols(_p(S), _p(out))
    return out


def inv2(a):
    a = _f(a)
    out = np.empty((2, 2), order="F")
    lib().orc_inv2(_p(a), _p(out))
    return out


def Traj_sim_general3(OdeTraj, Filter, t_correct, *, normals):
    """LNA_functional.cpp:771-817 with supplied normals (k*p, step-major)."""
    O = _f(OdeTraj)
    A, S = _f(Filter["A"]), _f(Filter["Sigma"])
    nrm = _d(normals)
    out = np.empty_like(O, order="F")
    ll = C.c_double()
    st = lib().orc_traj_sim_general3(_p(O), O.shape[0], O.shape[1] - 1, _p(A), _p(S), C.c_double(t_correct), _p(nrm),
                                     _p(out), C.byref(ll))
    if st:
        raise ValueError("chol(): decomposition failed")
    return {"SimuTraj": out, "loglike": ll.value}


def Traj_sim_ezG2(initial, times, param, gridsize, x_r, x_i, t_correct, transP="changepoint", model="SIR",
                  transX="standard", *, normals):
    """LNA_functional.cpp:907-975"""
    cfg = Cfg(x_r, x_i, len(param), model, transX)
    initial, times, param, nrm = _d(initial), _d(times), _d(param), _d(normals)
    nrow = len(times) // gridsize + 1
    out = np.empty((nrow, cfg.p + 1), order="F")
    ll = C.c_double()
    st = lib().orc_traj_sim_ezG2(cfg.ref, _p(initial), _p(times), len(times), _p(param), int(gridsize),
                                 C.c_double(t_correct), _p(nrm), _p(out), C.byref(ll))
    if st:
        raise ValueError("chol(): decomposition failed")
    return {"SimuTraj": out, "loglike": ll.value}


def ESlice_general2(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False,
                    model="SIR", transX="standard", *, normals, uniforms):
    """LNA_functional.cpp:1783-1857 (volz=True)"""
    if not volz:
        raise NotImplementedError("ESlice_general2(volz=False): Kingman branch is not restated; reached only with likelihood != 'volz', which no BASELINE config uses")
    F, O = _f(f_cur), _f(OdeTraj)
    A, S = _f(FTs["A"]), _f(FTs["Sigma"])
    ini = init if isinstance(init, Init) else Init(init)
    b, nrm, unf = _d(betaN), _d(normals), _d(uniforms)
    Index = (C.c_int * 2)(0, 1 if model == "SIR" else 2)
    out = np.empty_like(F, order="F")
    nu, ne = C.c_int(), C.c_int()
    st = lib().orc_eslice_general2(ini.ref, _p(F), F.shape[0], F.shape[1] - 1, _p(O), _p(A), _p(S), _p(b),
                                   C.c_double(t_correct), Index, TRANSX[transX], _p(nrm), _p(unf), C.byref(nu),
                                   C.byref(ne), _p(out))
    return {"newTraj": out, "status": st, "n_uniforms": nu.value, "n_evals": ne.value}


# ---- seasonal SIRS (src/SIRS_period.cpp) -------------------------------------------------------------

def ODE2(initial, t, param, period):
    initial, t, param = _d(initial), _d(t), _d(param)
    out = np.empty((len(t), 4), order="F")
    lib().orc_sirs_ode2(_p(initial), _p(t), len(t), _p(param), C.c_double(period), _p(out))
    return out


def SIRS_KOM_Filter(OdeTraj, param, gridsize, period):
    O, param = _f(OdeTraj), _d(param)
    k = (O.shape[0] - 1) // gridsize
    A, S = _cubes(3, k)
    lib().orc_sirs_kom_filter(_p(O), O.shape[0], _p(param), int(gridsize), C.c_double(period), _p(A), _p(S))
    return {"A": A, "Sigma": S}


def log_like_trajSIRS(SdeTraj, OdeTraj, Filter, gridsize, t_correct):
    X, O = _f(SdeTraj), _f(OdeTraj)
    A, S = _f(Filter["A"]), _f(Filter["Sigma"])
    ll = C.c_double()
    lib().orc_sirs_centred(0, _p(X), None, _p(O), O.shape[0], _p(A), _p(S), C.c_double(t_correct), None, C.byref(ll))
    return ll.value


def Traj_sim_SIRS(initial, OdeTraj, Filter, t_correct, *, normals):
    O, initial, z = _f(OdeTraj), _d(initial), _d(normals)
    A, S = _f(Filter["A"]), _f(Filter["Sigma"])
    out = np.empty_like(O, order="F")
    ll = C.c_double()
    st = lib().orc_sirs_centred(1, _p(z), _p(initial), _p(O), O.shape[0], _p(A), _p(S), C.c_double(t_correct), _p(out),
                                C.byref(ll))
    if st:
        raise ValueError("chol(): decomposition failed")
    return {"SimuTraj": out, "loglike": ll.value}


# ---- legacy copies of the pipeline: src/LNA_NS_general.cpp, class Foo of src/generalLNA.cpp, ESlice_SIRS ----
# x_i has two entries there: (nch, nparam); param = (R0, gamma, ..., ch_1, ..).

def _xi2(x_i):
    arr = np.ascontiguousarray(np.asarray(x_i, dtype=np.float64).round().astype(np.int32)[:2])
    return arr, arr.ctypes.data_as(_ip)


def betafs(ts, param, x_r, x_i):
    """LNA_NS_general.cpp:25-47"""
    ts, param, x_r = _d(ts), _d(param), _d(x_r)
    xi, xip = _xi2(x_i)
    out = np.empty(len(ts))
    lib().orc_general_betafs(_p(ts), len(ts), _p(param), _p(x_r), xip, _p(out))
    return out


def betaf(t, param, x_r, x_i):
    """LNA_NS_general.cpp:6-21"""
    return float(betafs([t], param, x_r, x_i)[0])


def _general_point(mode, states, param, t, x_r, x_i):
    states, param, x_r = _d(states), _d(param), _d(x_r)
    xi, xip = _xi2(x_i)
    out = np.empty((2, 2), order="F") if mode else np.empty(2)
    lib().orc_general_point(mode, _p(states), _p(param), len(param), C.c_double(t), _p(x_r), xip, _p(out))
    return out


def ODE_general_one(states, param, t, x_r, x_i):
    """LNA_NS_general.cpp:52-62"""
    return _general_point(0, states, param, t, x_r, x_i)


def general_F(states, param, t, x_r, x_i):
    """LNA_NS_general.cpp:66-72"""
    return _general_point(1, states, param, t, x_r, x_i)


def General_ODE_rk45(initial, t, param, x_r, x_i):
    """LNA_NS_general.cpp:76-98"""
    initial, t, param, x_r = _d(initial), _d(t), _d(param), _d(x_r)
    xi, xip = _xi2(x_i)
    out = np.empty((len(t), 3), order="F")
    lib().orc_general_ode_rk45(_p(initial), _p(t), len(t), _p(param), len(param), _p(x_r), xip, _p(out))
    return out


def General_IntSigmaF(Traj_par, param, x_r, x_i):
    """LNA_NS_general.cpp:103-135 (the reference names the second entry "Simga")"""
    T, param, x_r = _f(Traj_par), _d(param), _d(x_r)
    xi, xip = _xi2(x_i)
    F, S = np.empty((2, 2), order="F"), np.empty((2, 2), order="F")
    lib().orc_general_intsigmaf(_p(T), T.shape[0], _p(param), len(param), _p(x_r), xip, _p(F), _p(S))
    return {"expF": F, "Simga": S}


def General_KOM_Filter(OdeTraj, param, gridsize, x_r, x_i):
    """LNA_NS_general.cpp:139-159"""
    O, param, x_r = _f(OdeTraj), _d(param), _d(x_r)
    xi, xip = _xi2(x_i)
    k = (O.shape[0] - 1) // gridsize
    A, S = _cubes(2, k)
    lib().orc_general_kom_filter(_p(O), O.shape[0], _p(param), len(param), int(gridsize), _p(x_r), xip, _p(A), _p(S))
    return {"A": A, "Sigma": S}


def _legacy_centred(mode, X, OdeTraj, Filter, t_correct):
    O = _f(OdeTraj)
    A, S = _f(Filter["A"]), _f(Filter["Sigma"])
    X = _f(X) if mode in (0, 2) else _d(X)
    out = np.empty_like(O, order="F")
    ll = C.c_double()
    st = lib().orc_legacy_centred(mode, _p(X), _p(O), O.shape[0], _p(A), _p(S), C.c_double(t_correct), _p(out), C.byref(ll))
    if st:
        raise ValueError("legacy centred call failed")
    return out, ll.value


def log_like_traj_general(SdeTraj, OdeTraj, Filter, gridsize, t_correct, _mode=0):
    """LNA_NS_general.cpp:164-200 (density through inv2)"""
    return _legacy_centred(_mode, SdeTraj, OdeTraj, Filter, t_correct)[1]


def Traj_sim_general(OdeTraj, Filter, t_correct, *, normals, _mode=1):
    """LNA_NS_general.cpp:205-250 with supplied normals (k*2, step-major)"""
    out, ll = _legacy_centred(_mode, normals, OdeTraj, Filter, t_correct)
    return {"SimuTraj": out, "loglike": ll}


def Traj_sim_ezG(initial, times, param, gridsize, x_r, x_i, t_correct, *, normals):
    """LNA_NS_general.cpp:257-314"""
    initial, times, param, x_r, nrm = _d(initial), _d(times), _d(param), _d(x_r), _d(normals)
    xi, xip = _xi2(x_i)
    nrow = len(times) // gridsize + 1
    out = np.empty((nrow, 3), order="F")
    ll = C.c_double()
    st = lib().orc_traj_sim_ezG(_p(initial), _p(times), len(times), _p(param), len(param), int(gridsize), _p(x_r), xip,
                                C.c_double(t_correct), _p(nrm), _p(out), C.byref(ll))
    if st:
        raise ValueError("Traj_sim_ezG failed")
    return {"SimuTraj": out, "loglike": ll.value}


def volz_loglik_nh(init, f1, betaN, t_correct, gridsize=1):
    """SIR_phylodyn.cpp:517-560 (f1 = a LOG trajectory)"""
    ini = init if isinstance(init, Init) else Init(init)
    f1, b = _f(f1), _d(betaN)
    return float(lib().orc_volz_loglik_nh(ini.ref, _p(f1), f1.shape[0], _p(b), C.c_double(t_correct)))


def _eslice_legacy(kind, f_cur, OdeTraj, FTs, state, init, betaN_or_params4, t_correct, lam, reps, volz, normals, uniforms):
    F, O = _f(f_cur), _f(OdeTraj)
    A, S = _f(FTs["A"]), _f(FTs["Sigma"])
    ini = init if isinstance(init, Init) else Init(init)
    b, nrm, unf, state = _d(betaN_or_params4), _d(normals), _d(uniforms), _d(state)
    out = np.empty_like(F, order="F")
    nu, ne = C.c_int(), C.c_int()
    st = lib().orc_eslice_legacy(ini.ref, kind, _p(F), F.shape[0], F.shape[1] - 1, _p(O), _p(A), _p(S), _p(state), _p(b),
                                 C.c_double(t_correct), C.c_double(lam), int(reps), int(bool(volz)), _p(nrm), _p(unf),
                                 C.byref(nu), C.byref(ne), _p(out))
    if st == 1:
        raise ValueError("chol(): decomposition failed")
    return {"newTraj": out, "status": st, "n_uniforms": nu.value, "n_evals": ne.value}


def ESlice_general(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False, *,
                   normals, uniforms, _kind=0):
    """LNA_NS_general.cpp:320-394; normals reps x k*2, uniforms consumed contiguously over the repetitions"""
    return _eslice_legacy(_kind, f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam, reps, volz, normals, uniforms)


def ESlice_SIRS(f_cur, OdeTraj, FTs, state, init, params4, t_correct, lam=10.0, reps=1, gridsize=100, volz=True, *,
                normals, uniforms):
    """SIRS_period.cpp:313-402; params4 = (beta, A, period, N)"""
    return _eslice_legacy(1, f_cur, OdeTraj, FTs, state, init, params4, t_correct, lam, reps, volz, normals, uniforms)


class Foo:
    """class Foo of src/generalLNA.cpp:9-372 (m = 3 reactions, p = 2 species)"""

    def __init__(self, A_pre, A_post, param, x_i, x_r):
        self.A_pre, self.A_post = _f(A_pre), _f(A_post)
        self.param, self.x_i, self.x_r = _d(param), np.asarray(x_i), _d(x_r)
        self.m, self.p, self.N = self.A_pre.shape[0], self.A_pre.shape[1], float(self.x_r[0])

    def _point(self, mode, states, shape):
        out = np.empty(shape, order="F")
        lib().orc_foo_point(mode, _p(self.A_pre), _p(self.A_post), _p(self.param), C.c_double(self.N), _p(_d(states)), _p(out))
        return out

    def rate_h(self, states, t=0.0):
        return self._point(0, states, 3)

    def dh(self, states, t=0.0):
        return self._point(1, states, (3, 2))

    def ODE_ones(self, states, t=0.0):
        return self._point(2, states, 2)

    def F_matrix(self, states, t=0.0):
        """A' dh: what F_vec (:92-94) computes before its arma::vec return type rejects it"""
        return self._point(3, states, (2, 2))

    def ODE_rk45(self, initial, t):
        initial, t = _d(initial), _d(t)
        out = np.empty((len(t), 3), order="F")
        lib().orc_foo_ode_rk45(_p(self.A_pre), _p(self.A_post), _p(self.param), C.c_double(self.N), _p(initial), _p(t),
                               len(t), _p(out))
        return out

    def LNA_log_like_traj(self, SdeTraj, OdeTraj, Filter, gridsize, t_correct):
        return log_like_traj_general(SdeTraj, OdeTraj, Filter, gridsize, t_correct, _mode=2 if _is_ref() else 0)

    def LNA_Traj_sim(self, OdeTraj, Filter, t_correct, *, normals):
        return Traj_sim_general(OdeTraj, Filter, t_correct, normals=normals, _mode=3 if _is_ref() else 1)

    def LNA_ESlice(self, f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False, *,
                   normals, uniforms):
        return ESlice_general(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam, reps, gridsize, volz, normals=normals,
                              uniforms=uniforms, _kind=2 if _is_ref() else 0)


def _is_ref():
    """True when these wrappers run over the reference's code (oracle/reference.py swaps `lib`): there the Foo methods are
    reached through the class (modes 2, 3 / kind 2 of the harness); the C oracle has one body for both spellings."""
    return getattr(lib, "__module__", __name__) != __name__


# ---- data simulator -------------------------------------------------------------------------------------

def volz_thin_sir(eff_pop, t_correct, samp_times, n_sampled, betaN, lower=1.0, *, ua, ub):
    """R/coal_simulation.R:11-70 on a pre-drawn stream indexed by proposal number (ua -> waiting times, ub -> thinning marks)"""
    E = np.asarray(eff_pop, dtype=np.float64)
    t, S, I, b = _d(E[:, 0]), _d(E[:, 1]), _d(E[:, 2]), _d(betaN)
    st, ua, ub = _d(samp_times), _d(ua), _d(ub)
    ns_arr = np.ascontiguousarray(np.asarray(n_sampled, dtype=np.float64).round().astype(np.int32))
    max_coal = int(ns_arr.sum())
    ct, lin = np.empty(max_coal), np.empty(max_coal, dtype=np.int32)
    nc, nd = C.c_int(), C.c_long()
    rc = lib().orc_volz_thin_sir(_p(t), _p(S), _p(I), _p(b), len(t), C.c_double(t_correct), _p(st), ns_arr.ctypes.data_as(_ip),
                                 len(st), C.c_double(lower), _p(ua), _p(ub), C.c_long(len(ua)), _p(ct),
                                 lin.ctypes.data_as(_ip), max_coal, C.byref(nc), C.byref(nd))
    return {"coal_times": ct[:nc.value].copy(), "lineages": lin[:nc.value].copy(), "samp_times": st,
            "n_sampled": ns_arr.astype(np.float64), "n_draws": nd.value, "status": rc}
