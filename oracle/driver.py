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
            "betaN": npl["betaN"], "LatentTraj": lat, "coalLog": cl, "logOrigin": 0.0,
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


def opts_from_golden(g, update_traj=True):
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    ess = [0, 1, 0, 0, 1, 1, 0]
    if not update_traj:
        ess[4] = 0
    return {"priorList": [prior[0][:2], prior[1][:2], prior[2][:2], prior[4][:2]], "gamma_prior": tuple(prior[2][:2]),
            "gamma_prop": float(proposal[2][0]), "ESS_vec": ess, "update_traj": update_traj, "hyper_noop_mh": True}
