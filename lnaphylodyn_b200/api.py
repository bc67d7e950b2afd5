"""Host-side mirror of the reference's Rcpp-exported operator API for the hot path.

R is not available in this image, so this Python layer plays the role of R/RcppExports.R: the
function names, argument order/meaning and error behaviour follow the reference
(R/RcppExports.R:27-181; src/LNA_functional.h:15-120), and every call goes through the C ABI of
liblnaphylodyn_b200.so (include/lnaphylodyn_b200.h) into hand-written sm_100a kernels.
Differences from the reference, all forced by the batched/replayable design:
  * every function also accepts a leading batch dimension (param of shape (B, npt), ...);
  * samplers take pre-drawn `normals=` / `uniforms=` instead of calling an RNG;
  * Lists become dicts; arma::cube (p,p,k) becomes an ndarray of shape (p,p,k) (Fortran order).
No CPU fallback: importing works anywhere, calling needs a CUDA device.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import LnaCfg, LnaInit, LnaError, check

MODELS = {"SIR": 0, "SEIR": 1, "SIRS_period": 2}
MODEL_P = {"SIR": 2, "SEIR": 3, "SIRS_period": 3}
SENTINEL = -10000000.0
ST_CHOL_FAIL, ST_NEG_STATE, ST_NAN, ST_CAP_HIT = 1, 2, 4, 8
ESLICE_MAX_UNIF = 22
CHOL_MSG = "Invalid input for cholesky decomposition., parametrization fails"  # LNA_functional.cpp:663


def _d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return None if a is None else a.ctypes.data


def _xi(x_i):
    return [int(round(float(v))) for v in np.asarray(x_i).ravel()]


class Problem:
    """Device-resident (config + genealogy) handle: wraps lna_problem (lna_problem_create)."""

    def __init__(self, times, x_r, x_i, gridsize, t_correct=0.0, init=None, model="SIR", transX="standard",
                 device=0):
        if model not in MODELS:
            raise LnaError(f"model {model!r} is not on the GPU path (SIR, SEIR)")
        if transX not in ("standard", "log"):
            raise LnaError(f"transX {transX!r}: must be \"standard\" or \"log\"")
        self.times, self.x_r = _d(times), _d(x_r)
        self.x_i = _xi(x_i)
        self.model, self.gridsize, self.t_correct, self.device = model, int(gridsize), float(t_correct), int(device)
        cfg = LnaCfg(MODELS[model], 1 if transX == "log" else 0, self.x_i[0], self.x_i[1], self.x_i[2], self.x_i[3], self.gridsize,
                     len(self.times), self.times.ctypes.data_as(_lib._dp), self.x_r.ctypes.data_as(_lib._dp),
                     self.t_correct)
        ini_ref = None
        if init is not None:
            g = init
            self._C, self._D, self._y = _d(g["C"]), _d(g["D"]), _d(g["y"])
            self._gridrep = _d(g["gridrep"])
            ng = int(np.asarray(g["ng"]).ravel()[0])
            self._ini = LnaInit(len(self._C), ng, self._C.ctypes.data_as(_lib._dp), self._D.ctypes.data_as(_lib._dp),
                                self._y.ctypes.data_as(_lib._dp), self._gridrep.ctypes.data_as(_lib._dp))
            ini_ref = C.byref(self._ini)
        L = _lib.lib()
        self.h = L.lna_problem_create(C.byref(cfg), ini_ref, self.device)
        if not self.h:
            raise LnaError("lna_problem_create: " + L.lna_last_error().decode())
        p, k, nrow, npt = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(L.lna_problem_dims(self.h, C.byref(p), C.byref(k), C.byref(nrow), C.byref(npt)))
        self.p, self.k, self.nrow, self.npt = p.value, k.value, nrow.value, npt.value
        self.has_init = init is not None
        self._fin = weakref.finalize(self, L.lna_problem_destroy, self.h)

    # -- helpers -----------------------------------------------------------------------------
    def launch_count(self):
        return int(_lib.lib().lna_problem_launch_count(self.h))

    def sync(self):
        check(_lib.lib().lna_problem_sync(self.h), "lna_problem_sync")

    def cell_sums(self):
        n0 = self._ini.ng - 1
        cd, yy, cnt = np.empty(n0), np.empty(n0), np.empty(n0)
        check(_lib.lib().lna_problem_cell_sums(self.h, _ptr(cd), _ptr(yy), _ptr(cnt)), "lna_problem_cell_sums")
        return cd, yy, cnt

    def _batch(self, a, per):
        """-> (array of shape (B, per) C-contiguous, was_single)."""
        a = _d(a)
        single = a.ndim == 1 or (a.ndim >= 1 and a.size == per and a.ndim <= 2 and a.shape != (1, per) and a.ndim != 2)
        if a.size == per and a.ndim <= 1:
            return a.reshape(1, per), True
        return a.reshape(-1, per), False

    def _mats(self, a, shape):
        """item matrices given in natural (row, col[, slice]) indexing -> (B, prod) in column-major item layout."""
        a = np.asarray(a, dtype=np.float64)
        nd = len(shape)
        single = a.ndim == nd
        if single:
            a = a[None]
        assert a.shape[1:] == tuple(shape), (a.shape, shape)
        B = a.shape[0]
        flat = np.ascontiguousarray(np.transpose(a, (0,) + tuple(range(nd, 0, -1)))).reshape(B, -1)
        return flat, single

    def _unmats(self, flat, shape, single):
        nd = len(shape)
        a = flat.reshape((flat.shape[0],) + tuple(shape[::-1]))
        a = np.transpose(a, (0,) + tuple(range(nd, 0, -1)))
        return a[0] if single else a

    # -- batched operators -------------------------------------------------------------------
    def full_loglik(self, param, initial, OriginTraj=None, want=("coalLog",)):
        """FULL evaluation(s).  `want` subset of coalLog, FT, Ode, betaN, LatentTraj."""
        P, single = self._batch(param, self.npt)
        I, _ = self._batch(initial, self.p)
        B = P.shape[0]
        if I.shape[0] == 1 and B > 1:
            I = np.ascontiguousarray(np.broadcast_to(I, (B, self.p)))
        E = None
        if OriginTraj is not None:
            E, _ = self._mats(OriginTraj, (self.k, self.p))
            if E.shape[0] == 1 and B > 1:
                E = np.ascontiguousarray(np.broadcast_to(E, (B, E.shape[1])))
        p, k, nrow = self.p, self.k, self.nrow
        cl = np.empty(B) if "coalLog" in want else None
        st = np.zeros(B, dtype=np.int32)
        A = np.empty((B, p * p * k)) if "FT" in want else None
        Lc = np.empty((B, p * p * k)) if "FT" in want else None
        ode = np.empty((B, nrow * (p + 1))) if "Ode" in want else None
        bn = np.empty((B, nrow)) if "betaN" in want else None
        lat = np.empty((B, nrow * (p + 1))) if "LatentTraj" in want else None
        check(_lib.lib().lna_full_loglik(self.h, B, _ptr(P), _ptr(I), _ptr(E), _ptr(cl), _ptr(st), _ptr(A), _ptr(Lc),
                                         _ptr(ode), _ptr(bn), _ptr(lat)), "lna_full_loglik")
        out = {"status": st[0] if single else st}
        if cl is not None:
            out["coalLog"] = float(cl[0]) if single else cl
        if A is not None:
            out["FT"] = {"A": self._unmats(A, (p, p, k), single), "Sigma": self._unmats(Lc, (p, p, k), single)}
        if ode is not None:
            out["Ode"] = self._unmats(ode, (nrow, p + 1), single)
        if bn is not None:
            out["betaN"] = bn[0] if single else bn
        if lat is not None:
            out["LatentTraj"] = self._unmats(lat, (nrow, p + 1), single)
        return out


    def full_loglik_shared(self, param, initial, OriginTraj, asynchronous=False):
        """Chain-shared FULL evaluations (lna_full_loglik_shared): param (n_chains, proposals, npt), initial (n_chains, p),
        OriginTraj (n_chains, k, p) -> coalLog, status of shape (n_chains, proposals).  Every proposal of a chain is
        evaluated on the chain's initial state and latent increments -- the slice loop's body, ESlice_par_General :1588-1640."""
        P = _d(param)
        if P.ndim != 3 or P.shape[2] != self.npt:
            raise LnaError(f"full_loglik_shared: param must be (n_chains, proposals, {self.npt})")
        nc, m = P.shape[0], P.shape[1]
        I = _d(initial).reshape(nc, self.p)
        E, _ = self._mats(np.asarray(OriginTraj, dtype=np.float64).reshape(nc, self.k, self.p), (self.k, self.p))
        cl, st = np.empty((nc, m)), np.zeros((nc, m), dtype=np.int32)
        L = _lib.lib()
        fn = L.lna_full_loglik_shared_async if asynchronous else L.lna_full_loglik_shared
        check(fn(self.h, nc, m, _ptr(P), _ptr(I), _ptr(E), _ptr(cl), _ptr(st)), "lna_full_loglik_shared")
        if asynchronous:
            self.sync()
        return {"coalLog": cl, "status": st}


_cache = {}


def problem_for(times, x_r, x_i, gridsize, t_correct=0.0, init=None, model="SIR", transX="standard", device=0):
    """Problem handles are cached by value so that the Rcpp-style free functions stay cheap."""
    times, x_r = _d(times), _d(x_r)
    ini_key = None
    if init is not None:
        ini_key = (_d(init["C"]).tobytes(), _d(init["D"]).tobytes(), _d(init["y"]).tobytes(),
                   _d(init["gridrep"]).tobytes(), int(np.asarray(init["ng"]).ravel()[0]))
    key = (times.tobytes(), x_r.tobytes(), tuple(_xi(x_i)), int(gridsize), float(t_correct), ini_key, model, transX,
           device)
    pb = _cache.get(key)
    if pb is None:
        if len(_cache) > 32:
            _cache.clear()
        pb = _cache[key] = Problem(times, x_r, x_i, gridsize, t_correct, init, model, transX, device)
    return pb


def _grid_problem(times_like, x_r, x_i, model, transX, nt_min=2):
    """Problem for calls that carry no time grid of their own (betaTs, param_transform)."""
    t = np.array([0.0, 1.0]) if times_like is None else _d(times_like)
    return problem_for(t, x_r, x_i, 1, 0.0, None, model, transX)


# ------------------------------------------------------------------------------------------------
# Rcpp-named functions (R/RcppExports.R).  Single-item calls return the reference's shapes.
# ------------------------------------------------------------------------------------------------

def betaTs(param, times, x_r, x_i):
    """LNA_functional.cpp:31-58"""
    param = _d(param)
    xi = _xi(x_i)
    pb = problem_for(np.array([0.0, 1.0]), x_r, [xi[0], max(xi[1], 2), xi[2], xi[3]], 1)
    P = np.zeros((1, pb.npt))
    P[0, :min(pb.npt, param.size)] = param[:pb.npt]
    times = _d(times)
    out = np.empty((1, len(times)))
    check(_lib.lib().lna_betaTs(pb.h, 1, _ptr(P), _ptr(times), len(times), _ptr(out)), "lna_betaTs")
    return out[0]


def param_transform(t, param, x_r, x_i):
    """LNA_functional.cpp:63-86"""
    param = _d(param)
    xi = _xi(x_i)
    pb = problem_for(np.array([0.0, 1.0]), x_r, [xi[0], max(xi[1], 2), xi[2], xi[3]], 1)
    P = np.zeros((1, pb.npt))
    P[0, :min(pb.npt, param.size)] = param[:pb.npt]
    tt = np.array([float(t)])
    out = np.empty((1, pb.npt))
    check(_lib.lib().lna_param_transform(pb.h, 1, _ptr(tt), _ptr(P), _ptr(out)), "lna_param_transform")
    res = param.copy()
    res[0] = out[0, 0]
    return res


def _pad_param(pb, param):
    """The reference tolerates extra trailing params (testthat 'Test SIR model' (a)): only the first
    nparam+nch entries are read by the integrator."""
    param = _d(param)
    if param.ndim == 1:
        param = param[None]
    B = param.shape[0]
    P = np.zeros((B, pb.npt))
    n = min(pb.npt, param.shape[1])
    P[:, :n] = param[:, :n]
    if param.shape[1] < pb.npt - 1:
        raise LnaError(f"param has {param.shape[1]} entries, need at least nparam+nch = {pb.npt - 1}")
    return P


def ODE_rk45(initial, t, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    """LNA_functional.cpp:455-491 -> (n x (p+1)) matrix, column 0 = t."""
    pb = problem_for(t, x_r, x_i, 1, 0.0, None, model, transX)
    P = _pad_param(pb, param)
    I = _d(initial).reshape(-1, pb.p)
    B = P.shape[0]
    out = np.empty((B, len(pb.times) * (pb.p + 1)))
    check(_lib.lib().lna_ode_rk45(pb.h, B, _ptr(I), _ptr(P), _ptr(out)), "lna_ode_rk45")
    res = pb._unmats(out, (len(pb.times), pb.p + 1), np.asarray(param).ndim == 1)
    return res


def _kf(OdeTraj, param, gridsize, x_r, x_i, model, transX, chol):
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    pb = problem_for(times, x_r, x_i, gridsize, 0.0, None, model, transX)
    Of, _ = pb._mats(O, (len(times), pb.p + 1))
    P = _pad_param(pb, param)
    B = Of.shape[0]
    cube = pb.p * pb.p * pb.k
    A, S, st = np.empty((B, cube)), np.empty((B, cube)), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_kf_param(pb.h, B, _ptr(Of), _ptr(P), int(chol), _ptr(A), _ptr(S), _ptr(st)), "lna_kf_param")
    if chol and np.any(st & ST_CHOL_FAIL):
        raise ValueError(CHOL_MSG)
    return {"A": pb._unmats(A, (pb.p, pb.p, pb.k), single), "Sigma": pb._unmats(S, (pb.p, pb.p, pb.k), single)}


def KF_param(OdeTraj, param, gridsize, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    """LNA_functional.cpp:603-634"""
    return _kf(OdeTraj, param, gridsize, x_r, x_i, model, transX, 0)


def KF_param_chol(OdeTraj, param, gridsize, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    """LNA_functional.cpp:638-670 (the `Sigma` slot holds L)."""
    return _kf(OdeTraj, param, gridsize, x_r, x_i, model, transX, 1)


def SigmaF(Traj_par, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    """LNA_functional.cpp:536-599: one LNA interval = KF_param with gridsize = nrow(Traj_par).
    The rows' own times are used (a virtual closing row is appended, it is never read)."""
    T = np.asarray(Traj_par, dtype=np.float64)
    g = T.shape[0]
    last = T[-1].copy()
    last[0] = T[-1, 0] + (T[1, 0] - T[0, 0])
    res = _kf(np.vstack([T, last]), param, g, x_r, x_i, model, transX, 0)
    return {"expF": res["A"][:, :, 0], "Sigma": res["Sigma"][:, :, 0]}


def Ode_Coarse_Slicer(Ode_thin, gridsize):
    """LNA_functional.cpp:1179-1188"""
    O = np.asarray(Ode_thin, dtype=np.float64)
    p = O.shape[1] - 1
    model = "SIR" if p == 2 else "SEIR"
    pb = problem_for(O[:, 0], [1.0], [0, 2 if p == 2 else 3, 0, 1], gridsize, 0.0, None, model)
    Of, _ = pb._mats(O, (O.shape[0], p + 1))
    out = np.empty((1, pb.nrow * (p + 1)))
    check(_lib.lib().lna_ode_coarse_slicer(pb.h, 1, _ptr(Of), _ptr(out)), "lna_ode_coarse_slicer")
    return pb._unmats(out, (pb.nrow, p + 1), True)


def New_Param_List(param, initial, gridsize, t, x_r, x_i, transP="changepoint", model="SIR", transX="standard"):
    """LNA_functional.cpp:1191-1208 -> {FT:{A,Sigma(=L)}, Ode, betaN}; raises on Cholesky failure."""
    pb = problem_for(t, x_r, x_i, gridsize, 0.0, None, model, transX)
    res = pb.full_loglik(_pad_param(pb, param) if np.asarray(param).ndim > 1 else _pad_param(pb, param)[0], initial,
                         None, want=("FT", "Ode", "betaN"))
    if np.any(np.asarray(res["status"]) & ST_CHOL_FAIL):
        raise ValueError(CHOL_MSG)
    return {"FT": res["FT"], "Ode": res["Ode"], "betaN": res["betaN"]}


def _traj_problem(OdeTraj, p):
    O = np.asarray(OdeTraj, dtype=np.float64)
    times = O[:, 0] if O.ndim == 2 else O[0, :, 0]
    model = "SIR" if p == 2 else "SEIR"
    return problem_for(times, [1.0], [0, 2 if p == 2 else 3, 0, 1], 1, 0.0, None, model)


def TransformTraj(OdeTraj, OriginLatent, Filter_NC):
    """LNA_functional.cpp:877-900"""
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    p = O.shape[-1] - 1
    pb = _traj_problem(O, p)
    nrow = O.shape[-2]
    Of, _ = pb._mats(O, (nrow, p + 1))
    Ef, _ = pb._mats(OriginLatent, (nrow - 1, p))
    Af, _ = pb._mats(Filter_NC["A"], (p, p, nrow - 1))
    Lf, _ = pb._mats(Filter_NC["Sigma"], (p, p, nrow - 1))
    out = np.empty_like(Of)
    check(_lib.lib().lna_transform_traj(pb.h, Of.shape[0], _ptr(Of), _ptr(Ef), _ptr(Af), _ptr(Lf), _ptr(out)),
          "lna_transform_traj")
    return pb._unmats(out, (nrow, p + 1), single)


def Traj_sim_general_noncentral(OdeTraj, Filter_NC, t_correct, *, normals):
    """LNA_functional.cpp:824-872 with supplied normals (k*p, step-major)."""
    O = np.asarray(OdeTraj, dtype=np.float64)
    p, nrow = O.shape[-1] - 1, O.shape[-2]
    model = "SIR" if p == 2 else "SEIR"
    pb = problem_for(O[:, 0], [1.0], [0, 2 if p == 2 else 3, 0, 1], 1, t_correct, None, model)
    Of, _ = pb._mats(O, (nrow, p + 1))
    Af, _ = pb._mats(Filter_NC["A"], (p, p, nrow - 1))
    Lf, _ = pb._mats(Filter_NC["Sigma"], (p, p, nrow - 1))
    nrm = _d(normals).reshape(1, -1)
    traj, origin = np.empty_like(Of), np.empty((1, (nrow - 1) * p))
    lm, lo = np.empty(1), np.empty(1)
    check(_lib.lib().lna_traj_sim_noncentral(pb.h, 1, _ptr(Of), _ptr(Af), _ptr(Lf), _ptr(nrm), _ptr(traj), _ptr(origin),
                                             _ptr(lm), _ptr(lo)), "lna_traj_sim_noncentral")
    return {"SimuTraj": pb._unmats(traj, (nrow, p + 1), True), "OriginTraj": pb._unmats(origin, (nrow - 1, p), True),
            "logMultiNorm": float(lm[0]), "logOrigin": float(lo[0])}


def volz_loglik_nh2(init, f1, betaN, t_correct, index, transX="standard", _fn="lna_volz_loglik_nh2"):
    """LNA_functional.cpp:1119-1176 (literal per-event kernel)."""
    F = np.asarray(f1, dtype=np.float64)
    single = F.ndim == 2
    p, nrow = F.shape[-1] - 1, F.shape[-2]
    times = F[:, 0] if single else F[0, :, 0]
    ix = _xi(index)
    model = "SIR" if p == 2 else "SEIR"
    pb = problem_for(times, [1.0], [0, 2 if p == 2 else 3, ix[0], ix[1]], 1, t_correct, init, model, transX)
    Ff, _ = pb._mats(F, (nrow, p + 1))
    b = _d(betaN).reshape(Ff.shape[0], nrow)
    ll, st = np.empty(Ff.shape[0]), np.zeros(Ff.shape[0], dtype=np.int32)
    check(getattr(_lib.lib(), _fn)(pb.h, Ff.shape[0], _ptr(Ff), _ptr(b), _ptr(ll), _ptr(st)), _fn)
    return float(ll[0]) if single else ll


def volz_loglik_nh3(init, f1, betaN, t_correct, index, transX="standard"):
    """LNA_functional.cpp:1058-1113: volz_loglik_nh2 without the NaN guard."""
    return volz_loglik_nh2(init, f1, betaN, t_correct, index, transX, _fn="lna_volz_loglik_nh3")


def log_like_traj_general_adjust(SdeTraj, OdeTraj, Filter_NC, gridsize, t_correct):
    """LNA_functional.cpp:746-767: the body is commented out in the reference; it returns 0."""
    return 0.0


def Traj_sim_ezG_NC(initial, times, param, gridsize, x_r, x_i, t_correct, transP="changepoint", model="SIR",
                    transX="standard", *, normals):
    """LNA_functional.cpp:980-1004 = New_Param_List's ODE/FT + Traj_sim_general_noncentral (normals supplied)."""
    npl = New_Param_List(param, initial, gridsize, times, x_r, x_i, transP, model, transX)
    return Traj_sim_general_noncentral(npl["Ode"], npl["FT"], t_correct, normals=normals)


def ODE_rk45_stop(initial, t, param, x_r, x_i, transP="changepoint", model="SIR", transX="standard", tol=5):
    """LNA_functional.cpp:495-530: first grid time at which state x_i[4] drops below `tol`
    (else t[n-1] + 0.1).  The trajectory comes from the device; the scan is on the host."""
    traj = ODE_rk45(initial, t, param, x_r, x_i, transP, model, transX)
    col = traj[1:, 1 + _xi(x_i)[3]]
    hit = np.nonzero(col < tol)[0]
    tt = np.asarray(t, dtype=np.float64)
    return float(tt[1 + hit[0]]) if len(hit) else float(tt[-1] + 0.1)


def _kingman(init, f1, t_correct, lam, col, take_log):
    F = np.asarray(f1, dtype=np.float64)
    single = F.ndim == 2
    p, nrow = F.shape[-1] - 1, F.shape[-2]
    times = F[:, 0] if single else F[0, :, 0]
    model = "SIR" if p == 2 else "SEIR"
    pb = problem_for(times, [1.0], [0, 2 if p == 2 else 3, 0, 1], 1, t_correct, init, model)
    Ff, _ = pb._mats(F, (nrow, p + 1))
    lamv = np.ascontiguousarray(np.broadcast_to(_d(lam).ravel(), (Ff.shape[0],)))
    ll = np.empty(Ff.shape[0])
    check(_lib.lib().lna_coal_loglik(pb.h, Ff.shape[0], _ptr(Ff), _ptr(lamv), int(col), int(take_log), _ptr(ll)),
          "lna_coal_loglik")
    return float(ll[0]) if single else ll


def coal_loglik3(init, f1, t_correct, lam, Index, transX="standard"):
    """LNA_functional.cpp:1016-1054"""
    return _kingman(init, f1, t_correct, lam, Index + 1, transX == "standard")


def coal_loglik(init, f1, t_correct, lam, gridsize=1, transX="standard"):
    """SIR_phylodyn.cpp:436-468 (f1[,3] already holds log I)"""
    return _kingman(init, f1, t_correct, lam, 2, False)


def Update_Param(param, initial, t, OriginTraj, x_r, x_i, init, gridsize, coal_log=0.0, prior_proposal_offset=0.0,
                 t_correct=0.0, transP="changepoint", model="SIR", transX="standard", volz=True, *, unif):
    """LNA_functional.cpp:1212-1243; `unif` replaces the single R::runif(0,1) draw."""
    pb = problem_for(t, x_r, x_i, gridsize, t_correct, init, model, transX)
    P, single = pb._batch(param, pb.npt)
    B = P.shape[0]
    I, _ = pb._batch(initial, pb.p)
    E, _ = pb._mats(OriginTraj, (pb.k, pb.p))
    bc = lambda a, n: np.ascontiguousarray(np.broadcast_to(_d(a).reshape(-1, n) if n > 1 else _d(a).ravel(),
                                                           (B, n) if n > 1 else (B,)))
    I, E = bc(I, pb.p), bc(E, pb.k * pb.p)
    cl0, off, u = bc(coal_log, 1), bc(prior_proposal_offset, 1), bc(unif, 1)
    p, k, nrow = pb.p, pb.k, pb.nrow
    acc, cl, st = np.zeros(B, dtype=np.int32), np.empty(B), np.zeros(B, dtype=np.int32)
    A, Lc = np.empty((B, p * p * k)), np.empty((B, p * p * k))
    ode, bn, lat = np.empty((B, nrow * (p + 1))), np.empty((B, nrow)), np.empty((B, nrow * (p + 1)))
    check(_lib.lib().lna_update_param(pb.h, B, _ptr(P), _ptr(I), _ptr(E), _ptr(cl0), _ptr(off), _ptr(u), _ptr(acc),
                                      _ptr(cl), _ptr(st), _ptr(A), _ptr(Lc), _ptr(ode), _ptr(bn), _ptr(lat)),
          "lna_update_param")
    if single:
        if st[0] & ST_CHOL_FAIL:
            raise ValueError(CHOL_MSG)
        if not acc[0]:
            return {"accept": False, "coalLog_proposed": float(cl[0])}
        return {"accept": True, "FT": {"A": pb._unmats(A, (p, p, k), True), "Sigma": pb._unmats(Lc, (p, p, k), True)},
                "Ode": pb._unmats(ode, (nrow, p + 1), True), "betaN": bn[0], "coalLog": float(cl[0]),
                "LatentTraj": pb._unmats(lat, (nrow, p + 1), True)}
    return {"accept": acc.astype(bool), "status": st, "coalLog": cl,
            "FT": {"A": pb._unmats(A, (p, p, k), False), "Sigma": pb._unmats(Lc, (p, p, k), False)},
            "Ode": pb._unmats(ode, (nrow, p + 1), False), "betaN": bn, "LatentTraj": pb._unmats(lat, (nrow, p + 1), False)}


# ------------------------------------------------------------------------------------------------
# elliptical slice operators (pre-drawn streams; SURVEY Appendix C gives the consumption order)
# ------------------------------------------------------------------------------------------------

def _pr8(priorList):
    return _d(np.concatenate([np.asarray(v, dtype=np.float64).ravel()[:2] for v in priorList]))


def _i32(a, n):
    v = np.ascontiguousarray(np.asarray(a, dtype=np.float64).round().astype(np.int32).ravel())
    assert v.size == n, (v.size, n)
    return v


def _unif22(u, B):
    """pad the per-call uniform stream to the fixed LNA_ESLICE_MAX_UNIF width"""
    u = _d(u)
    u = u.reshape(B, -1) if u.ndim > 1 or B == 1 else u.reshape(B, -1)
    out = np.full((B, ESLICE_MAX_UNIF), 0.5)
    n = min(u.shape[1], ESLICE_MAX_UNIF)
    out[:, :n] = u[:, :n]
    return out


def Param_Slice_update(param, x_r, x_i, theta, newChs, rho=1):
    """LNA_functional.cpp:1246-1253"""
    xi = _xi(x_i)
    pb = problem_for(np.array([0.0, 1.0]), x_r, xi, 1)
    P, single = pb._batch(param, pb.npt)
    B = P.shape[0]
    th = np.ascontiguousarray(np.broadcast_to(_d(theta).ravel(), (B,)))
    nc = _d(newChs).reshape(B, xi[0])
    out = np.empty_like(P)
    check(_lib.lib().lna_param_slice_update(pb.h, B, _ptr(P), _ptr(th), _ptr(nc), _ptr(out)), "lna_param_slice_update")
    return out[0] if single else out


def Param_Slice_update_all2(par_old, x_r, x_i, theta, newChs, ESS_vec, priorList):
    """LNA_functional.cpp:1268-1314"""
    xi = _xi(x_i)
    pb = problem_for(np.array([0.0, 1.0]), x_r, xi, 1)
    npar = 2 + pb.npt
    P, single = pb._batch(par_old, npar)
    B = P.shape[0]
    th = np.ascontiguousarray(np.broadcast_to(_d(theta).ravel(), (B,)))
    nc = _d(newChs).reshape(B, npar - 1)
    ev, pr = _i32(ESS_vec, 7), _pr8(priorList)
    out = np.empty_like(P)
    check(_lib.lib().lna_param_slice_update_all2(pb.h, B, _ptr(P), _ptr(th), _ptr(nc), _ptr(ev), _ptr(pr), _ptr(out)),
          "lna_param_slice_update_all2")
    return out[0] if single else out


def ESlice_par_General(par_old, t, OriginTraj, priorList, x_r, x_i, init, gridsize, ESS_vec, coal_log=0.0,
                       t_correct=0.0, transP="changepoint", model="SIR", transX="standard", volz=True, *,
                       normals, uniforms, spec_width=0):
    """LNA_functional.cpp:1532-1688.  normals: npar-1 per chain; uniforms: u, theta draws (<= 21)."""
    pb = problem_for(t, x_r, x_i, gridsize, t_correct, init, model, transX)
    npar = 2 + pb.npt
    P, single = pb._batch(par_old, npar)
    B = P.shape[0]
    E, _ = pb._mats(OriginTraj, (pb.k, pb.p))
    if E.shape[0] == 1 and B > 1:
        E = np.ascontiguousarray(np.broadcast_to(E, (B, E.shape[1])))
    cl0 = np.ascontiguousarray(np.broadcast_to(_d(coal_log).ravel(), (B,)))
    nrm = _d(normals).reshape(B, -1)[:, :npar - 1].copy()
    unf = _unif22(uniforms, B)
    ev, pr = _i32(ESS_vec, 7), _pr8(priorList)
    p, k, nrow = pb.p, pb.k, pb.nrow
    par_new = np.empty((B, npar))
    A, Lc = np.empty((B, p * p * k)), np.empty((B, p * p * k))
    ode, bn, lat = np.empty((B, nrow * (p + 1))), np.empty((B, nrow)), np.empty((B, nrow * (p + 1)))
    cl, lo = np.empty(B), np.empty(B)
    ne, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    org = np.empty((B, k * p))
    check(_lib.lib().lna_eslice_par_general_ex(pb.h, B, _ptr(P), _ptr(E), _ptr(pr), _ptr(ev), _ptr(cl0), _ptr(nrm),
                                               _ptr(unf), int(spec_width), _ptr(par_new), _ptr(A), _ptr(Lc), _ptr(ode),
                                               _ptr(bn), _ptr(lat), _ptr(cl), _ptr(lo), _ptr(ne), _ptr(st), _ptr(org)),
          "lna_eslice_par_general")
    if single and (st[0] & ST_CHOL_FAIL):
        raise ValueError(CHOL_MSG)
    pick = (lambda a: a[0]) if single else (lambda a: a)
    # OriginTraj = OriginTraj_new of the result list: the input, or its running self-mixture with ESS_vec[7] = 1 (:1660-1662)
    return {"betaN": pick(bn), "FT": {"A": pb._unmats(A, (p, p, k), single), "Sigma": pb._unmats(Lc, (p, p, k), single)},
            "OdeTraj": pb._unmats(ode, (nrow, p + 1), single), "par": pick(par_new),
            "LatentTraj": pb._unmats(lat, (nrow, p + 1), single), "CoalLog": pick(cl),
            "OriginTraj": pb._unmats(org, (k, p), single),
            "logOrigin": pick(lo), "n_evals": pick(ne), "status": pick(st)}


def ESlice_general_NC(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, coal_log=-99999999.0,
                      gridsize=100, volz=False, model="SIR", transX="standard", *, normals, uniforms):
    """LNA_functional.cpp:1690-1778.  normals: k*p per chain, column-major k x p.  volz=False: the Kingman branch
    (coal_loglik3 of LogTraj(newTraj) with scale `lam`, :1733-1760; NaN as soon as I < 1, which ends the loop)."""
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    p = O.shape[-1] - 1
    xi = [0, 2, 0, 1]
    pb = problem_for(times, [1.0], xi, 1, t_correct, init, model, transX)
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, p + 1))
    B = Of.shape[0]
    Ff, _ = pb._mats(f_cur, (k, p))
    Af, _ = pb._mats(FTs["A"], (p, p, k))
    Lf, _ = pb._mats(FTs["Sigma"], (p, p, k))
    bn = _d(betaN).reshape(B, nrow)
    cl0 = np.ascontiguousarray(np.broadcast_to(_d(coal_log).ravel(), (B,)))
    nrm = _d(normals).reshape(B, k * p)
    unf = _unif22(uniforms, B)
    lat, org = np.empty((B, nrow * (p + 1))), np.empty((B, k * p))
    lo, cl = np.empty(B), np.empty(B)
    ne, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    if volz:
        check(_lib.lib().lna_eslice_general_nc(pb.h, B, _ptr(Ff), _ptr(Of), _ptr(Af), _ptr(Lf), _ptr(bn), _ptr(cl0),
                                               _ptr(nrm), _ptr(unf), _ptr(lat), _ptr(org), _ptr(lo), _ptr(cl), _ptr(ne),
                                               _ptr(st)), "lna_eslice_general_nc")
    else:
        lamv = np.ascontiguousarray(np.broadcast_to(_d(lam).ravel(), (B,)))
        check(_lib.lib().lna_eslice_general_nc_kingman(pb.h, B, _ptr(Ff), _ptr(Of), _ptr(Af), _ptr(Lf), _ptr(bn), _ptr(cl0),
                                                       _ptr(lamv), _ptr(nrm), _ptr(unf), _ptr(lat), _ptr(org), _ptr(lo),
                                                       _ptr(cl), _ptr(ne), _ptr(st)), "lna_eslice_general_nc_kingman")
    pick = (lambda a: a[0]) if single else (lambda a: a)
    return {"LatentTraj": pb._unmats(lat, (nrow, p + 1), single), "OriginTraj": pb._unmats(org, (k, p), single),
            "logOrigin": pick(lo), "CoalLog": pick(cl), "n_evals": pick(ne), "status": pick(st)}


def ESlice_change_points(param, initial, t, OriginTraj, x_r, x_i, init, gridsize, coal_log=0.0, t_correct=0.0,
                         transP="changepoint", model="SIR", transX="standard", volz=True, *, normals, uniforms,
                         spec_width=0):
    """LNA_functional.cpp:1317-1409.  uniforms: u, theta_1, shrinks; normals: nch per chain."""
    pb = problem_for(t, x_r, x_i, gridsize, t_correct, init, model, transX)
    P, single = pb._batch(param, pb.npt)
    B = P.shape[0]
    I, _ = pb._batch(initial, pb.p)
    if I.shape[0] == 1 and B > 1:
        I = np.ascontiguousarray(np.broadcast_to(I, (B, pb.p)))
    E, _ = pb._mats(OriginTraj, (pb.k, pb.p))
    if E.shape[0] == 1 and B > 1:
        E = np.ascontiguousarray(np.broadcast_to(E, (B, E.shape[1])))
    cl0 = np.ascontiguousarray(np.broadcast_to(_d(coal_log).ravel(), (B,)))
    nrm = _d(normals).reshape(B, -1)[:, :pb.x_i[0]].copy()
    unf = _unif22(uniforms, B)
    p, k, nrow = pb.p, pb.k, pb.nrow
    pn = np.empty((B, pb.npt))
    A, Lc = np.empty((B, p * p * k)), np.empty((B, p * p * k))
    ode, bn, lat = np.empty((B, nrow * (p + 1))), np.empty((B, nrow)), np.empty((B, nrow * (p + 1)))
    cl, lcp = np.empty(B), np.empty(B)
    ne, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_eslice_change_points(pb.h, B, _ptr(P), _ptr(I), _ptr(E), _ptr(cl0), _ptr(nrm), _ptr(unf),
                                              int(spec_width), _ptr(pn), _ptr(A), _ptr(Lc), _ptr(ode), _ptr(bn),
                                              _ptr(lat), _ptr(cl), _ptr(lcp), _ptr(ne), _ptr(st)),
          "lna_eslice_change_points")
    if single and (st[0] & ST_CHOL_FAIL):
        raise ValueError(CHOL_MSG)
    pick = (lambda a: a[0]) if single else (lambda a: a)
    return {"betaN": pick(bn), "FT": {"A": pb._unmats(A, (p, p, k), single), "Sigma": pb._unmats(Lc, (p, p, k), single)},
            "OdeTraj": pb._unmats(ode, (nrow, p + 1), single), "param": pick(pn),
            "LatentTraj": pb._unmats(lat, (nrow, p + 1), single), "CoalLog": pick(cl), "LogChProb": pick(lcp),
            "n_evals": pick(ne), "status": pick(st)}


# ------------------------------------------------------------------------------------------------
# centred parametrisation (legacy entry points of the same engine)
# ------------------------------------------------------------------------------------------------

def _centred_problem(OdeTraj, t_correct, init=None, model=None):
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    p = O.shape[-1] - 1
    model = model or ("SIR" if p == 2 else "SEIR")
    pb = problem_for(times, [1.0], [0, 2 if p == 2 else 3, 0, 1 if p == 2 else 2], 1, t_correct, init, model)
    return pb, O, single, p


def log_like_traj_general2(SdeTraj, OdeTraj, Filter, gridsize, t_correct):
    """LNA_functional.cpp:675-722"""
    pb, O, single, p = _centred_problem(OdeTraj, t_correct)
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, p + 1))
    Xf, _ = pb._mats(SdeTraj, (nrow, p + 1))
    Af, _ = pb._mats(Filter["A"], (p, p, k))
    Sf, _ = pb._mats(Filter["Sigma"], (p, p, k))
    ll = np.empty(Of.shape[0])
    check(_lib.lib().lna_log_like_traj_general2(pb.h, Of.shape[0], _ptr(Xf), _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(ll)),
          "lna_log_like_traj_general2")
    return float(ll[0]) if single else ll


def Traj_sim_general3(OdeTraj, Filter, t_correct, *, normals):
    """LNA_functional.cpp:771-817 with supplied normals (k*p, step-major)."""
    pb, O, single, p = _centred_problem(OdeTraj, t_correct)
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, p + 1))
    B = Of.shape[0]
    Af, _ = pb._mats(Filter["A"], (p, p, k))
    Sf, _ = pb._mats(Filter["Sigma"], (p, p, k))
    z = _d(normals).reshape(B, k * p)
    traj, ll, st = np.empty_like(Of), np.empty(B), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_traj_sim_general3(pb.h, B, _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(z), _ptr(traj), _ptr(ll),
                                           _ptr(st)), "lna_traj_sim_general3")
    if single and st[0] & ST_CHOL_FAIL:
        raise ValueError("chol(): decomposition failed")
    return {"SimuTraj": pb._unmats(traj, (nrow, p + 1), single), "loglike": float(ll[0]) if single else ll}


def Traj_sim_ezG2(initial, times, param, gridsize, x_r, x_i, t_correct, transP="changepoint", model="SIR",
                  transX="standard", *, normals):
    """LNA_functional.cpp:907-975 = ODE_rk45 + KF_param + coarse rows + the loop of Traj_sim_general3."""
    ode = ODE_rk45(initial, times, param, x_r, x_i, transP, model, transX)
    kf = KF_param(ode, param, gridsize, x_r, x_i, transP, model, transX)
    return Traj_sim_general3(Ode_Coarse_Slicer(ode, gridsize), kf, t_correct, normals=normals)


def ESlice_general2(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False,
                    model="SIR", transX="standard", *, normals, uniforms):
    """LNA_functional.cpp:1783-1857 (volz=True): centred elliptical slice on the trajectory; returns newTraj."""
    if not volz:
        raise LnaError("ESlice_general2(volz=False): the Kingman branch (LogTraj + coal_loglik3) is not built; reached only with likelihood != \"volz\", which no BASELINE config uses")
    pb, O, single, p = _centred_problem(OdeTraj, t_correct, init, model)
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, p + 1))
    B = Of.shape[0]
    Ff, _ = pb._mats(f_cur, (nrow, p + 1))
    Af, _ = pb._mats(FTs["A"], (p, p, k))
    Sf, _ = pb._mats(FTs["Sigma"], (p, p, k))
    bn = _d(betaN).reshape(B, nrow)
    z = _d(normals).reshape(B, k * p)
    unf = _unif22(uniforms, B)
    out = np.empty_like(Of)
    ne, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_eslice_general2(pb.h, B, _ptr(Ff), _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(bn), _ptr(z), _ptr(unf),
                                         _ptr(out), _ptr(ne), _ptr(st)), "lna_eslice_general2")
    pick = (lambda a: a[0]) if single else (lambda a: a)
    return {"newTraj": pb._unmats(out, (nrow, p + 1), single), "n_evals": pick(ne), "status": pick(st)}


# ------------------------------------------------------------------------------------------------
# seasonal SIRS model variant (src/SIRS_period.cpp): ODE / LNA moments / trajectory density
# ------------------------------------------------------------------------------------------------

def _sirs_problem(times, period, gridsize=1):
    return problem_for(times, [float(period)], [0, 4, 0, 1], gridsize, 0.0, None, "SIRS_period")


def ODE2(initial, t, param, period):
    """SIRS_period.cpp:174-193: seasonal SIRS mean trajectory, param = (beta, gamma, mu, alpha)."""
    pb = _sirs_problem(t, period)
    P = _d(param).reshape(-1, 4)
    I = _d(initial).reshape(-1, 3)
    out = np.empty((P.shape[0], len(pb.times) * 4))
    check(_lib.lib().lna_ode_rk45(pb.h, P.shape[0], _ptr(I), _ptr(P), _ptr(out)), "lna_ode_rk45")
    return pb._unmats(out, (len(pb.times), 4), np.asarray(param).ndim == 1)


def SIRS_KOM_Filter(OdeTraj, param, gridsize, period):
    """SIRS_period.cpp:147-168 (+ SIRS_IntSigma :50-79): {A, Sigma} per LNA interval."""
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    pb = _sirs_problem(O[:, 0] if single else O[0, :, 0], period, gridsize)
    Of, _ = pb._mats(O, (len(pb.times), 4))
    P = _d(param).reshape(-1, 4)
    B = Of.shape[0]
    A, S, st = np.empty((B, 9 * pb.k)), np.empty((B, 9 * pb.k)), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_kf_param(pb.h, B, _ptr(Of), _ptr(P), 0, _ptr(A), _ptr(S), _ptr(st)), "lna_kf_param")
    return {"A": pb._unmats(A, (3, 3, pb.k), single), "Sigma": pb._unmats(S, (3, 3, pb.k), single)}


def SIRS_loglik_sweep(param, initial, times, SdeTraj, gridsize, t_correct, period, want=("loglik",), device=0):
    """The seasonal-SIRS likelihood sweep over parameter draws (BASELINE configs[3]), fused: per draw
        log_like_trajSIRS(SdeTraj, Ode_Coarse_Slicer(ODE2(initial, times, param, period), gridsize),
                          SIRS_KOM_Filter(ODE2(...), param, gridsize, period), gridsize, t_correct)
    (SIRS_period.cpp:174-193, :147-168, :82-118) in one kernel launch, one thread per draw.  param (B, 4); initial (3,) shared
    or (B, 3); SdeTraj (nrow, 4) shared by the sweep, (B, nrow, 4) per draw, or (B / m, nrow, 4) per m consecutive draws; None
    with want = ("A", "Sigma", "Ode") for the moments only."""
    pb = problem_for(times, [float(period)], [0, 4, 0, 1], gridsize, t_correct, None, "SIRS_period", device=device)
    P = _d(param).reshape(-1, 4)
    B = P.shape[0]
    I = _d(initial)
    shared_init = I.ndim == 1 or I.shape[0] == 1
    I = I.reshape(-1, 3)
    nrow, k = pb.nrow, pb.k
    X, group = None, 1
    if SdeTraj is not None:
        X, one = pb._mats(SdeTraj, (nrow, 4))
        if one or X.shape[0] == 1:
            group = 0
        elif X.shape[0] != B:
            if B % X.shape[0]:
                raise LnaError("SIRS_loglik_sweep: the number of draws must be a multiple of the number of trajectories")
            group = B // X.shape[0]
    ll = np.empty(B) if "loglik" in want else None
    if ll is not None and X is None:
        raise LnaError("SIRS_loglik_sweep: a log-likelihood needs SdeTraj")
    st = np.zeros(B, dtype=np.int32)
    A = np.empty((B, 9 * k)) if "A" in want else None
    S = np.empty((B, 9 * k)) if "Sigma" in want else None
    O = np.empty((B, nrow * 4)) if "Ode" in want else None
    check(_lib.lib().lna_sirs_loglik(pb.h, B, _ptr(P), _ptr(I), int(shared_init), _ptr(X), int(group), _ptr(ll), _ptr(st),
                                     _ptr(A), _ptr(S), _ptr(O)), "lna_sirs_loglik")
    out = {"status": st}
    if ll is not None:
        out["loglik"] = ll
    if A is not None:
        out["A"] = pb._unmats(A, (3, 3, k), False)
    if S is not None:
        out["Sigma"] = pb._unmats(S, (3, 3, k), False)
    if O is not None:
        out["Ode"] = pb._unmats(O, (nrow, 4), False)
    return out


def Traj_sim_SIRS(initial, OdeTraj, Filter, t_correct, *, normals):
    """SIRS_period.cpp:197-240 with supplied normals (k*3, step-major).  A failed chol(Sigma + 1e-13 I) -- the covariance has
    rank 2, the outcome hangs on its last rounding -- raises like the reference for a single call and is reported in
    `status` (ST_CHOL_FAIL) per item for a batch."""
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    pb = problem_for(times, [1.0], [0, 3, 0, 2], 1, t_correct, None, "SEIR")  # any 3-compartment problem on this grid
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, 4))
    B = Of.shape[0]
    Af, _ = pb._mats(Filter["A"], (3, 3, k))
    Sf, _ = pb._mats(Filter["Sigma"], (3, 3, k))
    I = np.ascontiguousarray(np.broadcast_to(_d(initial).reshape(-1, 3), (B, 3)))
    z = _d(normals).reshape(B, k * 3)
    T, ll, st = np.empty((B, nrow * 4)), np.empty(B), np.zeros(B, dtype=np.int32)
    check(_lib.lib().lna_traj_sim_sirs(pb.h, B, _ptr(I), _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(z), _ptr(T), _ptr(ll), _ptr(st)),
          "lna_traj_sim_sirs")
    if single and (st[0] & ST_CHOL_FAIL):
        raise ValueError("chol(): decomposition failed")
    return {"SimuTraj": pb._unmats(T, (nrow, 4), single), "loglike": float(ll[0]) if single else ll,
            "status": st[0] if single else st}


def log_like_trajSIRS(SdeTraj, OdeTraj, Filter, gridsize, t_correct):
    """SIRS_period.cpp:82-118"""
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    pb = problem_for(times, [1.0], [0, 3, 0, 2], 1, t_correct, None, "SEIR")  # any 3-compartment problem on this grid
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, 4))
    Xf, _ = pb._mats(SdeTraj, (nrow, 4))
    Af, _ = pb._mats(Filter["A"], (3, 3, k))
    Sf, _ = pb._mats(Filter["Sigma"], (3, 3, k))
    ll = np.empty(Of.shape[0])
    check(_lib.lib().lna_log_like_traj_sirs(pb.h, Of.shape[0], _ptr(Xf), _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(ll)),
          "lna_log_like_traj_sirs")
    return float(ll[0]) if single else ll


# ------------------------------------------------------------------------------------------------
# legacy copies of the pipeline: src/LNA_NS_general.cpp, class Foo of src/generalLNA.cpp, ESlice_SIRS.
# x_i has two entries there, (nch, nparam); param = (R0, gamma, .., ch_1 ..).  The integrator / moment functions are
# arithmetic twins of the current ones on the SIR model and run on the same kernels.
# ------------------------------------------------------------------------------------------------

def _legacy_xi(x_i):
    xi = [int(round(float(v))) for v in np.asarray(x_i).ravel()[:2]]
    return [xi[0], xi[1], 0, 1]


def _legacy_param(param, x_i):
    """(R0, gamma, .., ch_1..ch_nch) -> the current layout's param with a trailing (unused) hyper-parameter"""
    xi = _legacy_xi(x_i)
    param = _d(param)
    return np.append(param[:xi[0] + xi[1]], 1.0)


def betafs(ts, param, x_r, x_i):
    """LNA_NS_general.cpp:25-47"""
    return betaTs(_legacy_param(param, x_i), ts, x_r, _legacy_xi(x_i))


def betaf(t, param, x_r, x_i):
    """LNA_NS_general.cpp:6-21"""
    return float(betafs([t], param, x_r, x_i)[0])


_SIR_PRE = np.array([1.0, 0.0, 0.0, 1.0, 1.0, 0.0])    # 3 x 2 column-major: S + I -> 2 I, I -> 0, (no third reaction)
_SIR_POST = np.array([0.0, 0.0, 0.0, 2.0, 0.0, 0.0])


def _foo_call(mode, A_pre, A_post, param3, N, states, t=None):
    pb = problem_for(np.array([0.0, 1.0]), [1.0], [0, 2, 0, 1], 1)

    def colmajor(A):  # 3 x 2 matrix (or its 6 column-major entries)
        A = np.asarray(A, dtype=np.float64)
        return np.ascontiguousarray(A.reshape(-1, order="F") if A.ndim == 2 else A)

    ap, aq = colmajor(A_pre), colmajor(A_post)
    par, Nv, x = _d(param3).reshape(1, 3), np.array([float(N)]), _d(states).reshape(1, 2)
    n = 0 if t is None else len(t)
    per = {0: 3, 1: 6, 2: 2, 3: 4, 4: n * 3, 5: 2, 6: 4}[mode]
    out = np.empty((1, per))
    tt = None if t is None else _d(t)
    check(_lib.lib().lna_foo(pb.h, 1, mode, n, _ptr(ap), _ptr(aq), _ptr(par), _ptr(Nv), _ptr(x),
                             None if tt is None else _ptr(tt), _ptr(out)), "lna_foo")
    return out[0]


def ODE_general_one(states, param, t, x_r, x_i):
    """LNA_NS_general.cpp:52-62"""
    th1 = betaf(t, param, x_r, x_i)
    return _foo_call(5, _SIR_PRE, _SIR_POST, [th1, _d(param)[1], 0.0], 1.0, states)


def general_F(states, param, t, x_r, x_i):
    """LNA_NS_general.cpp:66-72"""
    th1 = betaf(t, param, x_r, x_i)
    return _foo_call(6, _SIR_PRE, _SIR_POST, [th1, _d(param)[1], 0.0], 1.0, states).reshape(2, 2, order="F")


def General_ODE_rk45(initial, t, param, x_r, x_i):
    """LNA_NS_general.cpp:76-98"""
    return ODE_rk45(initial, t, _legacy_param(param, x_i), x_r, _legacy_xi(x_i))


def General_IntSigmaF(Traj_par, param, x_r, x_i):
    """LNA_NS_general.cpp:103-135 (the reference spells the second entry "Simga")"""
    r = SigmaF(Traj_par, _legacy_param(param, x_i), x_r, _legacy_xi(x_i))
    return {"expF": r["expF"], "Simga": r["Sigma"]}


def General_KOM_Filter(OdeTraj, param, gridsize, x_r, x_i):
    """LNA_NS_general.cpp:139-159"""
    return KF_param(OdeTraj, _legacy_param(param, x_i), gridsize, x_r, _legacy_xi(x_i))


def log_like_traj_general(SdeTraj, OdeTraj, Filter, gridsize, t_correct):
    """LNA_NS_general.cpp:164-200 (density through inv2)"""
    pb, O, single, p = _centred_problem(OdeTraj, t_correct)
    if p != 2:
        raise LnaError("log_like_traj_general: the legacy copies are SIR-only (p = 2)")
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, 3))
    Xf, _ = pb._mats(SdeTraj, (nrow, 3))
    Af, _ = pb._mats(Filter["A"], (2, 2, k))
    Sf, _ = pb._mats(Filter["Sigma"], (2, 2, k))
    ll = np.empty(Of.shape[0])
    check(_lib.lib().lna_log_like_traj_general(pb.h, Of.shape[0], _ptr(Xf), _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(ll)),
          "lna_log_like_traj_general")
    return float(ll[0]) if single else ll


def Traj_sim_general(OdeTraj, Filter, t_correct, *, normals):
    """LNA_NS_general.cpp:205-250 with supplied normals (k*2, step-major)"""
    pb, O, single, p = _centred_problem(OdeTraj, t_correct)
    if p != 2:
        raise LnaError("Traj_sim_general: the legacy copies are SIR-only (p = 2)")
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, 3))
    B = Of.shape[0]
    Af, _ = pb._mats(Filter["A"], (2, 2, k))
    Sf, _ = pb._mats(Filter["Sigma"], (2, 2, k))
    z = _d(normals).reshape(B, k * 2)
    traj, ll = np.empty_like(Of), np.empty(B)
    check(_lib.lib().lna_traj_sim_general(pb.h, B, _ptr(Of), _ptr(Af), _ptr(Sf), _ptr(z), _ptr(traj), _ptr(ll)),
          "lna_traj_sim_general")
    return {"SimuTraj": pb._unmats(traj, (nrow, 3), single), "loglike": float(ll[0]) if single else ll}


def Traj_sim_ezG(initial, times, param, gridsize, x_r, x_i, t_correct, *, normals):
    """LNA_NS_general.cpp:257-314 = General_ODE_rk45 + General_KOM_Filter + coarse rows + the loop of Traj_sim_general"""
    ode = General_ODE_rk45(initial, times, param, x_r, x_i)
    kf = General_KOM_Filter(ode, param, gridsize, x_r, x_i)
    return Traj_sim_general(Ode_Coarse_Slicer(ode, gridsize), kf, t_correct, normals=normals)


def volz_loglik_nh(init, f1, betaN, t_correct, gridsize=1):
    """SIR_phylodyn.cpp:517-560 (f1 = a LOG trajectory, nrow x 3)"""
    F = np.asarray(f1, dtype=np.float64)
    single = F.ndim == 2
    times = F[:, 0] if single else F[0, :, 0]
    pb = problem_for(times, [1.0], [0, 2, 0, 1], 1, t_correct, init, "SIR")
    Ff, _ = pb._mats(F, (pb.nrow, 3))
    B = Ff.shape[0]
    bn = _d(betaN).reshape(B, pb.nrow)
    ll = np.empty(B)
    check(_lib.lib().lna_volz_loglik_nh(pb.h, B, _ptr(Ff), _ptr(bn), _ptr(ll)), "lna_volz_loglik_nh")
    return float(ll[0]) if single else ll


def _eslice_legacy(kind, f_cur, OdeTraj, FTs, state, init, betaN_or_params4, t_correct, lam, reps, volz, normals, uniforms):
    O = np.asarray(OdeTraj, dtype=np.float64)
    single = O.ndim == 2
    times = O[:, 0] if single else O[0, :, 0]
    p = 3 if kind == 1 else 2
    if O.shape[-1] != p + 1:
        raise LnaError(f"trajectory has {O.shape[-1] - 1} compartments, this sampler works on {p}")
    pb = problem_for(times, [1.0], [0, 3, 0, 2] if p == 3 else [0, 2, 0, 1], 1, t_correct, init, "SEIR" if p == 3 else "SIR")
    nrow, k = pb.nrow, pb.k
    Of, _ = pb._mats(O, (nrow, p + 1))
    B = Of.shape[0]
    Ff, _ = pb._mats(f_cur, (nrow, p + 1))
    Af, _ = pb._mats(FTs["A"], (p, p, k))
    Sf, _ = pb._mats(FTs["Sigma"], (p, p, k))
    q = _d(betaN_or_params4).reshape(B, 4 if kind == 1 else nrow)
    st0 = None if kind != 1 else np.ascontiguousarray(np.broadcast_to(_d(state).reshape(-1, 3), (B, 3)))
    lamv = np.ascontiguousarray(np.broadcast_to(_d(lam).reshape(-1), (B,)))
    # normals arrive per item as reps x k*p (the order the reference consumes them in); the C ABI wants repetition-major
    z = np.ascontiguousarray(_d(normals).reshape(B, reps, k * p).transpose(1, 0, 2))
    unf = np.zeros((B, reps * 22))
    u = _d(uniforms).reshape(B, -1)
    unf[:, :min(u.shape[1], reps * 22)] = u[:, :reps * 22]
    out = np.empty_like(Of)
    ne, nu, st = (np.zeros(B, dtype=np.int32) for _ in range(3))
    check(_lib.lib().lna_eslice_legacy(pb.h, B, kind, int(reps), int(bool(volz)), _ptr(Ff), _ptr(Of), _ptr(Af), _ptr(Sf),
                                       None if st0 is None else _ptr(st0), _ptr(q), _ptr(lamv), _ptr(z), _ptr(unf), _ptr(out),
                                       _ptr(ne), _ptr(nu), _ptr(st)), "lna_eslice_legacy")
    if single and st[0] & ST_CHOL_FAIL:
        raise ValueError("chol(): decomposition failed")
    pick = (lambda a: a[0]) if single else (lambda a: a)
    return {"newTraj": pb._unmats(out, (nrow, p + 1), single), "n_evals": pick(ne), "n_uniforms": pick(nu), "status": pick(st)}


def ESlice_general(f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False, *,
                   normals, uniforms):
    """LNA_NS_general.cpp:320-394; normals reps x k*2 per item, uniforms consumed contiguously over the repetitions"""
    return _eslice_legacy(0, f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam, reps, volz, normals, uniforms)


def ESlice_SIRS(f_cur, OdeTraj, FTs, state, init, params4, t_correct, lam=10.0, reps=1, gridsize=100, volz=True, *,
                normals, uniforms):
    """SIRS_period.cpp:313-402; params4 = (beta, A, period, N)"""
    return _eslice_legacy(1, f_cur, OdeTraj, FTs, state, init, params4, t_correct, lam, reps, volz, normals, uniforms)


class Foo:
    """class Foo of src/generalLNA.cpp:9-372 (exposed to R through RCPP_MODULE(mod_Foo)): a reaction network with m = 3
    reactions on p = 2 species, hazards hard-wired to (param0 param1 / N S I, param1 I, param2 N)."""

    def __init__(self, A_pre, A_post, param, x_i, x_r):
        self.A_pre, self.A_post = np.asarray(A_pre, dtype=np.float64), np.asarray(A_post, dtype=np.float64)
        if self.A_pre.shape != (3, 2) or self.A_post.shape != (3, 2):
            raise LnaError("Foo: rate_h / dh are written for 3 reactions on 2 species (generalLNA.cpp:76-90)")
        self.param, self.x_i, self.x_r = _d(param), np.asarray(x_i), _d(x_r)
        self.m, self.p, self.N = 3, 2, float(self.x_r[0])

    # get / set functions (:38-47)
    def get_A_pre(self):
        return self.A_pre

    def get_A_post(self):
        return self.A_post

    def get_A(self):
        return self.A_post - self.A_pre

    def get_param(self):
        return self.param

    def set_param(self, new_param):
        self.param = _d(new_param)

    def set_x_i(self, new_x_i):
        self.x_i = np.asarray(new_x_i)

    def set_x_r(self, new_x_r):
        self.x_r = _d(new_x_r)   # (N is NOT refreshed: the reference copies it in the constructor only, :35)

    def transform_param(self, t=0.0):
        """:50-55"""
        h = self.rate_h([1.0, 1.0], t)  # param_new(0) = h(0) at S = I = 1
        res = self.param.copy()
        res[0] = h[0]
        return res

    def _call(self, mode, states, t=None):
        return _foo_call(mode, self.A_pre, self.A_post, self.param[:3], self.N, states, t)

    def rate_h(self, states, t=0.0):
        return self._call(0, states)

    def dh(self, states, t=0.0):
        return self._call(1, states).reshape(3, 2, order="F")

    def ODE_ones(self, states, t=0.0):
        return self._call(2, states)

    def F_matrix(self, states, t=0.0):
        """A' dh (2 x 2): the product F_vec (:92-94) forms"""
        return self._call(3, states).reshape(2, 2, order="F")

    def F_vec(self, states, t=0.0):
        raise LnaError("Foo::F_vec (generalLNA.cpp:92-94) returns the 2 x 2 product A' dh through an arma::vec: the reference "
                       "throws here; F_matrix() returns the product")

    def ODE_rk45(self, initial, t):
        t = _d(t)
        return self._call(4, initial, t).reshape(len(t), 3, order="F")

    def IntSigmaF(self, Traj_par):
        raise LnaError("Foo::IntSigmaF (generalLNA.cpp:124-156) cannot run as written: it calls F_vec and multiplies by an "
                       "empty local `A` (:130, 146); use General_IntSigmaF")

    def KOM_Filter_MX(self, OdeTraj, gridsize):
        return self.IntSigmaF(None)

    def LNA_Traj_sim_ez(self, initial, times, gridsize, t_correct):
        return self.IntSigmaF(None)

    def LNA_log_like_traj(self, SdeTraj, OdeTraj, Filter, gridsize, t_correct):
        """:178-214"""
        return log_like_traj_general(SdeTraj, OdeTraj, Filter, gridsize, t_correct)

    def LNA_Traj_sim(self, OdeTraj, Filter, t_correct, *, normals):
        """:217-262"""
        return Traj_sim_general(OdeTraj, Filter, t_correct, normals=normals)

    def LNA_ESlice(self, f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam=10.0, reps=1, gridsize=100, volz=False, *,
                   normals, uniforms):
        """:296-372"""
        return _eslice_legacy(2, f_cur, OdeTraj, FTs, state, init, betaN, t_correct, lam, reps, volz, normals, uniforms)
