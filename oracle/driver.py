"""CPU restatement of the reference's R-level MCMC iteration (TEST INFRASTRUCTURE, see lna_oracle.h).

`eslice_iteration` follows General_MCMC_with_ESlice's loop body (R/LNA_chpt_slice.R:775-847) as the
Changpoint_sim / Ebola vignettes configure it:
  Update_Param_each_Norm(parIdlist = list(c = 4, d = nch + 5), isLoglist = (1, 0), priorIdlist = (3, 5),
                         proposeIdlist = priorIdlist, hyper = F)          R/LNA_chpt_slice.R:169-265
  update_Par_ESlice_combine -> ESlice_par_General                         R/LNA_chpt_slice.R:456-472
  updateTraj_general_NC     -> ESlice_general_NC + volz_loglik_nh2        R/LNA_chpt_slice.R:372-415
composed from the C oracle's functions, with pre-drawn streams laid out like the product's
lna_chains_step: normals = [npar-1 for the parameter ESS | k*p for the trajectory ESS],
uniforms = [gamma RW draw, gamma MH accept, hyper MH accept, 22 for the parameter ESS, 22 for the trajectory ESS].
"""
import math

import numpy as np

from . import oracle as orc

M_LN_SQRT_2PI = 0.918938533204672741780329736406


def dnorm_log(x, mu, sigma):
    z = (x - mu) / sigma
    return -(M_LN_SQRT_2PI + 0.5 * z * z + np.log(sigma))


def init_state(g, par, OriginTraj):
    """MCMC_initialize_general with control$traj given (R/general_change_point.R:507-536)."""
    par = np.asarray(par, dtype=np.float64).copy()
    npl = orc.New_Param_List(par[2:], par[:2], g.gridsize, g.times, g.x_r, g.x_i)
    lat = orc.TransformTraj(npl["Ode"], OriginTraj, npl["FT"])
    cl = orc.volz_loglik_nh2(g.init, lat, npl["betaN"], g.t_correct, g.x_i[2:4])
    return {"par": par, "OriginTraj": np.array(OriginTraj, dtype=np.float64), "FT": npl["FT"], "Ode": npl["Ode"],
            "betaN": npl["betaN"], "LatentTraj": lat, "coalLog": cl,
            "logOrigin": -float(np.sum(np.asarray(OriginTraj, dtype=np.float64) ** 2)) / 2,  # general_change_point.R:530
            "n_full": 0, "n_cheap": 0, "n_mh_accept": 0}


def _mh(st, g, par_new, offset, unif):
    try:
        r = orc.Update_Param(par_new[2:], par_new[:2], g.times, st["OriginTraj"], g.x_r, g.x_i, g.init, g.gridsize,
                             st["coalLog"], offset, g.t_correct, unif=unif)
    except ValueError:  # tryCatch(...) returns MCMC_obj unchanged
        st["n_full"] += 1
        return
    st["n_full"] += 1
    if r["accept"]:
        st["par"] = par_new
        st["Ode"], st["betaN"], st["coalLog"], st["LatentTraj"], st["FT"] = (r["Ode"], r["betaN"], r["coalLog"],
                                                                             r["LatentTraj"], r["FT"])
        st["n_mh_accept"] += 1


def eslice_iteration(st, g, opts, normals, uniforms):
    npar = len(st["par"])
    nch = npar - 5
    # --- entry c: log gamma' = log gamma + runif(-po, po)  (:217-218)
    par_new = st["par"].copy()
    raw = np.log(par_new[3])
    po = opts["gamma_prop"]
    raw_new = raw + (-po + (po - (-po)) * uniforms[0])
    newdiff = dnorm_log(raw_new, *opts["gamma_prior"]) - dnorm_log(raw, *opts["gamma_prior"])
    par_new[3] = np.exp(raw_new)
    _mh(st, g, par_new, newdiff, uniforms[1])
    # --- entry d (hyper = F): no proposal, but ch <- exp(log(ch) * hyper / hyper)  (:238-241)
    if opts.get("hyper_noop_mh", True):
        par_new = st["par"].copy()
        hy = par_new[npar - 1]
        par_new[4:4 + nch] = np.exp(np.log(par_new[4:4 + nch]) * hy / hy)
        _mh(st, g, par_new, 0.0, uniforms[2])
    # --- parameter ESS
    r = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                               g.gridsize, opts["ESS_vec"], st["coalLog"], g.t_correct,
                               normals=normals[:npar - 1], uniforms=uniforms[3:25])
    st["n_full"] += r["n_evals"]
    if not (r["status"] & 1):
        st["par"], st["LatentTraj"], st["betaN"], st["FT"] = r["par"], r["LatentTraj"], r["betaN"], r["FT"]
        st["coalLog"], st["Ode"] = r["CoalLog"], r["OdeTraj"]  # logOrigin is NOT refreshed here (:456-472)
    st["last_par_evals"] = r["n_evals"]
    # --- ESS_vec2: a second parameter ESS INSTEAD of the trajectory ESS (R/LNA_chpt_slice.R:826-833); it takes the first
    #     npar-1 normals and the 22 uniforms of the trajectory update's slots
    if opts.get("update_traj", True) and opts.get("ESS_vec2") is not None:
        r = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                                   g.gridsize, opts["ESS_vec2"], st["coalLog"], g.t_correct,
                                   normals=normals[npar - 1:2 * (npar - 1)], uniforms=uniforms[25:47])
        st["n_full"] += r["n_evals"]
        if not (r["status"] & 1):
            st["par"], st["LatentTraj"], st["betaN"], st["FT"] = r["par"], r["LatentTraj"], r["betaN"], r["FT"]
            st["coalLog"], st["Ode"] = r["CoalLog"], r["OdeTraj"]
        return st
    # --- trajectory ESS
    if opts.get("update_traj", True):
        betaN = orc.betaTs(st["par"][2:npar - 1], st["LatentTraj"][:, 0], g.x_r, g.x_i)
        t = orc.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, betaN, g.t_correct,
                                  1.0, st["coalLog"], g.gridsize, True, "SIR", normals=normals[npar - 1:],
                                  uniforms=uniforms[25:47])
        new_cl = orc.volz_loglik_nh2(g.init, t["LatentTraj"], betaN, g.t_correct, g.x_i[2:4])
        st["coalLog"], st["LatentTraj"], st["OriginTraj"], st["logOrigin"] = (new_cl, t["LatentTraj"], t["OriginTraj"],
                                                                              t["logOrigin"])
        st["n_cheap"] += t["n_evals"]
        st["last_traj_evals"] = t["n_evals"]
    return st


def eslice_iteration_general(st, g, opts, normals, uniforms):
    """General_MCMC_with_ESlice's loop body with an ARBITRARY Update_Param_each_Norm entry list (R/LNA_chpt_slice.R:169-265),
    e.g. the Ebola vignette's parIdlist = (1, 3, nch + 5).  opts["entries"]: (0-based par index or -1 for the hyper entry,
    isLog, (prior mean, sd), proposal half-width); opts["hyper"]: the R call's `hyper` flag.  Streams as the product's general
    layout: uniforms [2e, 2e+1] = (walk, accept) of entry e, [8:30] parameter ESS, [30:52] trajectory ESS."""
    npar = len(st["par"])
    nch = npar - 5
    for e, (idx, is_log, pr, po) in enumerate(opts["entries"]):
        par_new = st["par"].copy()
        if idx >= 0:  # priorId in 1:4 (:211-217)
            raw = math.log(par_new[idx]) if is_log else float(par_new[idx])
            raw_new = raw + (-po + (po - (-po)) * uniforms[2 * e])
            off = dnorm_log(raw_new, pr[0], pr[1]) - dnorm_log(raw, pr[0], pr[1])
            par_new[idx] = math.exp(raw_new) if is_log else raw_new
        else:  # the hyper entry (priorId 5)
            hy = float(par_new[npar - 1])
            if opts.get("hyper", False):  # :219-227
                u = -po + (po - (-po)) * uniforms[2 * e]
                hy_new = hy * math.exp(u)
                ln, lo = math.log(hy_new), math.log(hy)
                off = (dnorm_log(ln, pr[0], pr[1]) - ln) - u - (dnorm_log(lo, pr[0], pr[1]) - lo)
            else:
                hy_new, off = hy, 0.0
            par_new[4:4 + nch] = np.exp(np.log(par_new[4:4 + nch]) * hy / hy_new)  # :238-241
            par_new[npar - 1] = hy_new
        _mh(st, g, par_new, off, uniforms[2 * e + 1])
    r = orc.ESlice_par_General(st["par"], g.times, st["OriginTraj"], opts["priorList"], g.x_r, g.x_i, g.init,
                               g.gridsize, opts["ESS_vec"], st["coalLog"], g.t_correct,
                               normals=normals[:npar - 1], uniforms=uniforms[8:30])
    st["n_full"] += r["n_evals"]
    if not (r["status"] & 1):
        st["par"], st["LatentTraj"], st["betaN"], st["FT"] = r["par"], r["LatentTraj"], r["betaN"], r["FT"]
        st["coalLog"], st["Ode"] = r["CoalLog"], r["OdeTraj"]
    if opts.get("update_traj", True):
        betaN = orc.betaTs(st["par"][2:npar - 1], st["LatentTraj"][:, 0], g.x_r, g.x_i)
        t = orc.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, betaN, g.t_correct,
                                  1.0, st["coalLog"], g.gridsize, True, "SIR", normals=normals[npar - 1:],
                                  uniforms=uniforms[30:52])
        new_cl = orc.volz_loglik_nh2(g.init, t["LatentTraj"], betaN, g.t_correct, g.x_i[2:4])
        st["coalLog"], st["LatentTraj"], st["OriginTraj"], st["logOrigin"] = (new_cl, t["LatentTraj"], t["OriginTraj"],
                                                                              t["logOrigin"])
        st["n_cheap"] += t["n_evals"]
    return st


def opts_from_golden(g, update_traj=True):
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    ess = [0, 1, 0, 0, 1, 1, 0]
    if not update_traj:
        ess[4] = 0
    return {"priorList": [prior[0][:2], prior[1][:2], prior[2][:2], prior[4][:2]], "gamma_prior": tuple(prior[2][:2]),
            "gamma_prop": float(proposal[2][0]), "ESS_vec": ess, "update_traj": update_traj, "hyper_noop_mh": True}


# ------------------------------------------------------------------------------------------------
# General_MCMC2 iteration body (R/LNA_chpt_slice.R:576-691), stages i < burn1 and burn1 <= i < warmup2
# ------------------------------------------------------------------------------------------------
import math


def dunif_log(x, a, b):
    return -math.log(b - a) if a <= x <= b else -math.inf


def dgamma_log(x, shape, rate):
    """closed form of dgamma(x, shape, rate, log = TRUE) (R goes through dpois_raw; unpinned)"""
    return shape * math.log(rate) - math.lgamma(shape) + (shape - 1.0) * math.log(x) - rate * x


def mcmc2_iteration(st, g, opts, normals, uniforms):
    """uniforms: [2e, 2e+1] MH entry e (walk, accept); [8, 9] hyper; [10:32] change-point ESS; [32:54]
    trajectory ESS.  normals: [0:nch] change-point ESS, [nch:] trajectory ESS."""
    npar = len(st["par"])
    nch = npar - 5
    # Update_Param_each (R/LNA_chpt_slice.R:266-365)
    for e, (idx, is_log, kind, pr, po) in enumerate(opts["entries"]):
        par_new = st["par"].copy()
        raw = math.log(par_new[idx]) if is_log else par_new[idx]
        raw_new = raw + (-po + (po - (-po)) * uniforms[2 * e])
        if kind == 0:
            newdiff = dnorm_log(raw_new, pr[0], pr[1]) - dnorm_log(raw, pr[0], pr[1])
        else:
            newdiff = dunif_log(raw_new, pr[0], pr[1]) - dunif_log(raw, pr[0], pr[1])
        par_new[idx] = math.exp(raw_new) if is_log else raw_new
        _mh(st, g, par_new, newdiff, uniforms[2 * e + 1])
    # update_hyper (R/general_change_point.R:281-313)
    if opts.get("update_hyper", True):
        par_new = st["par"].copy()
        hs = par_new[npar - 1]
        hp = opts["hyper_prop"]
        u = -hp + (hp - (-hp)) * uniforms[8]
        hs_new = hs * math.exp(u)
        par_new[npar - 1] = hs_new
        par_new[4:4 + nch] = np.exp(np.log(par_new[4:4 + nch]) * hs / hs_new)
        off = dgamma_log(hs_new, *opts["hyper_prior"]) - dgamma_log(hs, *opts["hyper_prior"]) + u
        _mh(st, g, par_new, off, uniforms[9])
    # update_ChangePoint_ESlice (R/LNA_chpt_slice.R:419-440)
    if opts.get("update_chpt", True):
        r = orc.ESlice_change_points(st["par"][2:], st["par"][:2], g.times, st["OriginTraj"], g.x_r, g.x_i, g.init,
                                     g.gridsize, st["coalLog"], g.t_correct, normals=normals[:nch],
                                     uniforms=uniforms[10:32])
        st["n_full"] += r["n_evals"]
        if not (r["status"] & 1):
            st["par"] = np.concatenate([st["par"][:2], r["param"]])
            st["LatentTraj"], st["betaN"], st["FT"] = r["LatentTraj"], r["betaN"], r["FT"]
            st["coalLog"], st["Ode"], st["chprobs"] = r["CoalLog"], r["OdeTraj"], r["LogChProb"]
    if opts.get("update_traj", True):
        betaN = orc.betaTs(st["par"][2:npar - 1], st["LatentTraj"][:, 0], g.x_r, g.x_i)
        t = orc.ESlice_general_NC(st["OriginTraj"], st["Ode"], st["FT"], st["par"][:2], g.init, betaN, g.t_correct,
                                  1.0, st["coalLog"], g.gridsize, True, "SIR", normals=normals[nch:],
                                  uniforms=uniforms[32:54])
        new_cl = orc.volz_loglik_nh2(g.init, t["LatentTraj"], betaN, g.t_correct, g.x_i[2:4])
        st["coalLog"], st["LatentTraj"], st["OriginTraj"], st["logOrigin"] = (new_cl, t["LatentTraj"], t["OriginTraj"],
                                                                              t["logOrigin"])
        st["n_cheap"] += t["n_evals"]
    return st


def mcmc2_joint_iteration(st, g, opts, T, normals, uniforms):
    """General_MCMC2's stage i >= warmup2 (R/LNA_chpt_slice.R:655-683): update_Param_joint(method = "admcmc") (:9-88), then
    the change-point and trajectory slice updates.  `T` (d x d): RawTransParam_new = RawTransParam + T z, what
    MASS::mvrnorm(1, RawTransParam, PCOV) computes with T = eS$vectors diag(sqrt(pmax(ev, 0))) (third-party: the
    eigenvector signs are not pinned, so T is an input here).  normals: [0:nch] change-point ESS, [nch:nch+k*p] trajectory
    ESS, [nch+k*p : +d] the joint proposal; uniforms: [1] accept of the joint step, [10:32], [32:54] as mcmc2_iteration."""
    npar = len(st["par"])
    nch = npar - 5
    kp = len(normals) - nch - 4
    z = normals[nch + kp:nch + kp + len(opts["joint"])]
    par_new = st["par"].copy()
    raw, raw_new, off = [], [], 0.0
    for i, (idx, is_log, kind, pr) in enumerate(opts["joint"]):
        j = npar - 1 if idx == 4 else idx
        r = math.log(par_new[j]) if is_log else float(par_new[j])
        rn = r
        for q in range(len(opts["joint"])):
            rn = rn + T[i][q] * z[q]
        if kind == 0:
            off += dnorm_log(rn, pr[0], pr[1]) - dnorm_log(r, pr[0], pr[1])
        elif kind == 1:
            off += dunif_log(rn, pr[0], pr[1]) - dunif_log(r, pr[0], pr[1])
        else:
            off += dgamma_log(rn, pr[0], pr[1]) - dgamma_log(r, pr[0], pr[1])
        raw.append(r)
        raw_new.append(rn)
    for i, (idx, is_log, kind, pr) in enumerate(opts["joint"]):
        j = npar - 1 if idx == 4 else idx
        if idx == 4:  # :65-68, with the OLD and NEW hyper-parameter
            par_new[4:4 + nch] = np.exp(np.log(par_new[4:4 + nch]) * raw[i] / raw_new[i])
        par_new[j] = math.exp(raw_new[i]) if is_log else raw_new[i]
    _mh(st, g, par_new, off, uniforms[1])
    sub = dict(opts, entries=(), update_hyper=False)
    return mcmc2_iteration(st, g, sub, normals[:nch + kp], uniforms)


def mcmc2_opts_from_golden(g, burn=False):
    """Options of the constant_sim vignette call (vignettes/constant_sim.Rmd:87-96)."""
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    ent = []
    for idx, is_log, pid in ((1, 1, 1), (2, 0, 2), (3, 1, 3)):
        ent.append((idx, is_log, 1 if pid == 2 else 0, tuple(prior[pid - 1][:2]), float(proposal[pid - 1][0])))
    return {"entries": ent, "update_hyper": not burn, "hyper_prop": float(proposal[4][0]),
            "hyper_prior": tuple(prior[4][:2]), "update_chpt": True, "update_traj": not burn}
