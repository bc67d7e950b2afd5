"""R-level drivers of the reference, batched: `General_MCMC_with_ESlice` and `General_MCMC2`
(R/LNA_chpt_slice.R:508-693, 723-849) with `MCMC_setup_general` / `MCMC_initialize_general`
(R/general_change_point.R:416-567) for `n_chains` independent chains on one GPU.

Same argument names, same meaning and the same result list as the R functions (`par`, `Trajectory`, `l`, `l1`,
`MCMC_setting`, `MCMC_obj`), with a leading chain axis when `n_chains > 1`.  The iteration bodies run on the
device through the resident-chain C ABI (`lna_chains_step*`, see chains.py); this module is the run loop around
them: set-up, initial state, stage switching (`burn1`, `warmup2`), thinning and storage.

What is NOT here, and raises instead of guessing: likelihood other than "volz", models other than "SIR" (the
reference's samplers are hard-wired to p = 2, LNA_functional.cpp:1456-1462), vector-valued Metropolis entries.
General_MCMC2's adaptive joint random-walk stage (`i >= warmup2`) draws its proposal as
MASS::mvrnorm does (eigen-decomposition of the covariance), with LAPACK's eigenvector signs: same proposal distribution, the
realisation for given normals is not pinned (MASS is outside the reference tree).
The random numbers are Philox streams drawn on the device (seed, chain, iteration) -- R's generator is not
reproduced, so a run is distributionally, not bitwise, the R run; per-iteration arithmetic on given
streams is pinned by tests/test_gpu_eslice.py and tests/test_gpu_vs_reference.py.
"""
import ctypes as C

import numpy as np

from . import _lib, api, chains as chm, genealogy
from ._lib import LnaError, check

_PRIOR_ORDER = ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")
_PROP_ORDER = ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")


def MCMC_setup_general(coal_obs, times, t_correct, N, gridsize=50, niter=1000, burn=500, thin=5, changetime=(),
                       DEMS=("E", "I"), prior=None, proposal=None, control=None, likelihood="volz", model="SIR",
                       Index=(0, 1), nparam=2, PCOV=None):
    """R/general_change_point.R:416-435.  `coal_obs`: dict with samp_times, n_sampled, coal_times."""
    times = np.asarray(times, dtype=np.float64)
    gridset = np.arange(0, len(times), int(gridsize))
    grid = times[gridset]
    Init = genealogy.coal_lik_init(coal_obs["samp_times"], coal_obs["n_sampled"], coal_obs["coal_times"], grid)
    if t_correct is None:
        t_correct = float(np.max(coal_obs["coal_times"]))
    changetime = np.asarray(changetime, dtype=np.float64).ravel()
    return {"Init": Init, "times": times, "t_correct": float(t_correct), "x_r": np.concatenate([[float(N)], changetime]),
            "gridsize": int(gridsize), "gridset": gridset + 1, "niter": int(niter), "burn": int(burn), "thin": int(thin),
            "x_i": [len(changetime), int(nparam), int(Index[0]), int(Index[1])], "prior": dict(prior or {}),
            "proposal": dict(proposal or {}), "control": dict(control or {}), "p": len(DEMS), "reps": 1,
            "likelihood": likelihood, "model": model, "PCOV": PCOV}


def _problem(setting, device=0):
    return api.problem_for(setting["times"], setting["x_r"], setting["x_i"], setting["gridsize"], setting["t_correct"],
                           setting["Init"], model=setting["model"], device=device)


def MCMC_initialize_general(MCMC_setting, n_chains=1, rng=None, device=0, max_tries=50):
    """R/general_change_point.R:439-567 for `n_chains` chains: `control` values are used where given, the rest is
    drawn from the priors (numpy generator `rng`); a chain is redrawn while its initial coalescent likelihood is the
    -1e7 sentinel / NaN, like the R `while` loop.  Returns the batch as a dict of arrays with a leading chain axis."""
    s = MCMC_setting
    if s["model"] != "SIR" or s["p"] != 2:
        raise LnaError("MCMC_initialize_general: only model = 'SIR' (p = 2) has samplers on the GPU path")
    rng = rng or np.random.default_rng(0)
    ctl, pr = s["control"], s["prior"]
    nch, B = s["x_i"][0], int(n_chains)
    pb = _problem(s, device)
    k, p = pb.k, pb.p
    par = np.empty((B, 2 + 2 + nch + 1))
    eps = np.empty((B, k, p))
    todo = np.ones(B, dtype=bool)
    coal = np.full(B, np.nan)
    for _ in range(max_tries):
        n = int(todo.sum())
        if n == 0:
            break
        q = np.empty((n, par.shape[1]))
        q[:, 0] = s["x_r"][0]
        if ctl.get("alpha") is None:
            q[:, 1] = rng.lognormal(pr["pop_pr"][0], pr["pop_pr"][1], n)
        else:
            q[:, 1] = float(np.ravel(ctl["alpha"])[0]) * s["x_r"][0]
        q[:, 2] = rng.uniform(pr["R0_pr"][0], pr["R0_pr"][1], n) if ctl.get("R0") is None else float(ctl["R0"])
        q[:, 3] = (np.exp(rng.normal(pr["gamma_pr"][0], pr["gamma_pr"][1], n)) if ctl.get("gamma") is None
                   else float(ctl["gamma"]))
        hyper = (rng.gamma(pr["hyper_pr"][0], 1.0 / pr["hyper_pr"][1], n) if ctl.get("hyper") is None
                 else np.full(n, float(ctl["hyper"])))
        q[:, -1] = hyper
        if nch:
            q[:, 4:4 + nch] = (rng.lognormal(0.0, 1.0, (n, nch)) ** (1.0 / hyper)[:, None] if ctl.get("ch") is None
                               else np.asarray(ctl["ch"], dtype=np.float64).ravel()[None, :])
        e = (rng.standard_normal((n, k, p)) if ctl.get("traj") is None
             else np.broadcast_to(np.asarray(ctl["traj"], dtype=np.float64), (n, k, p)))
        res = pb.full_loglik(q[:, 2:], q[:, :2], e, want=("coalLog",))
        cl = np.atleast_1d(res["coalLog"])
        idx = np.flatnonzero(todo)
        ok = np.isfinite(cl) & (cl > -10000000.0)
        par[idx[ok]], eps[idx[ok]], coal[idx[ok]] = q[ok], e[ok], cl[ok]
        todo[idx[ok]] = False
        if ctl.get("traj") is not None and all(ctl.get(v) is not None for v in ("alpha", "R0", "gamma", "hyper", "ch")):
            if not ok.all():
                raise LnaError("MCMC_initialize_general: the fully specified control state has no finite likelihood")
    if todo.any():
        raise LnaError(f"MCMC_initialize_general: {int(todo.sum())} chains found no finite initial likelihood")
    ch = par[:, 4:4 + nch]
    chprobs = -0.5 * np.sum((np.log(ch) * par[:, -1:]) ** 2, axis=1)
    return {"par": par, "OriginTraj": eps, "coalLog": coal, "logOrigin": -0.5 * np.sum(eps * eps, axis=(1, 2)),
            "chprobs": chprobs, "p": p}


def _lists(d, order, what):
    try:
        return [np.asarray(d[k], dtype=np.float64).ravel() for k in order]
    except KeyError as e:
        raise LnaError(f"{what} list needs the entries {order} (missing {e})")


def _finish(ch, setting, rec, n_chains, squeeze):
    s = ch.get()
    obj = {"par": s["par"], "LatentTraj": s["LatentTraj"], "OriginTraj": s["OriginTraj"], "coalLog": s["coalLog"],
           "logOrigin": s["logOrigin"], "p": 2, "n_full_evals": s["n_full_evals"], "n_cheap_evals": s["n_cheap_evals"],
           "n_mh_accept": s["n_mh_accept"]}
    out = {"par": rec["par"], "Trajectory": rec["traj"], "l": rec["l"], "l1": rec["l1"], "l2": None, "l3": None,
           "MX": setting["PCOV"], "MCMC_setting": setting, "MCMC_obj": obj}
    if squeeze and n_chains == 1:
        for key in ("par", "Trajectory", "l", "l1"):
            out[key] = out[key][0]
        out["MCMC_obj"] = {k: (v[0] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == 1 else v)
                           for k, v in obj.items()}
    return out


def _record(ch, rec, i, thin, want_traj):
    """params[i,] = par; l[i] = logOrigin; l1[i] = coalLog; tjs[,,i/thin] = LatentTraj  (LNA_chpt_slice.R:839-845).
    par, l, l1 are recorded ON THE DEVICE (lna_chains_history_record: device-to-device copies, no host round trip per
    iteration) and fetched once by _fetch_history; only the thinned trajectories cross to the host as they are produced."""
    ch.history_record(i)
    if want_traj:
        rec["traj"][:, (i + 1) // thin - 1] = ch.latent()


def _begin_history(ch, niter):
    ch.history_begin(niter)


def _fetch_history(ch, rec):
    """The device-side records into rec["par"] (B, niter, npar), rec["l"], rec["l1"] (B, niter)."""
    ch.history_get(rec["par"], rec["l"], rec["l1"])


def _make_chains(setting, init, device, devices):
    """The resident batch: one GPU (`device`), or sharded over `devices` (an int = that many GPUs from 0, or a list) and
    driven by this process through lna_multi."""
    if devices is None:
        pb = _problem(setting, device)
        return pb, chm.Chains(pb, init["par"], init["OriginTraj"], free_running=True)
    from . import comm
    devs = list(range(int(devices))) if isinstance(devices, (int, np.integer)) else [int(d) for d in devices]
    mc = comm.MultiChains({"times": setting["times"], "x_r": setting["x_r"], "x_i": setting["x_i"],
                           "gridsize": setting["gridsize"], "t_correct": setting["t_correct"], "init": setting["Init"]},
                          init["par"], init["OriginTraj"], devs)
    return mc.pb, mc


def General_MCMC_with_ESlice(coal_obs, times, t_correct, N, gridsize=1000, niter=1000, burn=0, thin=5, changetime=(),
                             DEMS=("S", "I"), prior=None, proposal=None, control=None, ESS_vec=(1, 1, 1, 1, 1, 1, 1),
                             likelihood="volz", model="SIR", Index=(0, 1), nparam=2, method="seq", options=None,
                             ESS_vec2=None, verbose=False, *, n_chains=1, seed=0, device=0, devices=None, chain_offset=0,
                             squeeze=True, init_rng=None):
    """R/LNA_chpt_slice.R:723-849 for `n_chains` chains.  Iterations i < options["burn1"] run the Metropolis entries
    and the parameter slice update with ESS_vec[5] switched off (ESS_temp); later ones add the trajectory slice update.
    Returns {par (chains, niter, npar), Trajectory (chains, niter/thin, nrow, p+1), l, l1, MCMC_setting, MCMC_obj}."""
    if likelihood != "volz" or model != "SIR":
        raise LnaError("General_MCMC_with_ESlice: the GPU path has the volz likelihood with model = 'SIR' only")
    options = dict(options or {})
    nch = len(np.ravel(changetime))
    par_ids = [int(v) for v in dict(options.get("parIdlist") or {"c": 4, "d": nch + 5}).values()]
    is_logs = [int(v) for v in dict(options.get("isLoglist") or {"c": 1, "d": 0}).values()]
    pr_ids = [int(v) for v in dict(options.get("priorIdlist") or {"c": 3, "d": 5}).values()]
    hyper_flag = bool(options.get("hyper", False))
    if not (len(par_ids) == len(is_logs) == len(pr_ids)):
        raise LnaError("General_MCMC_with_ESlice: parameters do not match")  # check_list_eq, R/LNA_chpt_slice.R:175-185
    for pid_, prid_ in zip(par_ids, pr_ids):
        if not ((1 <= pid_ <= 4 and 1 <= prid_ <= 4) or (pid_ == nch + 5 and prid_ == 5)):
            raise LnaError("General_MCMC_with_ESlice: Update_Param_each_Norm entries are built for parId 1..4 (S0, I0, R0, gamma) "
                           "with priorId 1..4 and for the hyper-parameter (parId nch + 5, priorId 5)")
    # the vignettes' (gamma, no-op hyper) pair keeps its dedicated path: both entries in one latency period
    standard_entries = par_ids == [4, nch + 5] and is_logs == [1, 0] and pr_ids == [3, 5] and not hyper_flag
    setting = MCMC_setup_general(coal_obs, times, t_correct, N, gridsize, niter, burn, thin, changetime, DEMS, prior,
                                 proposal, control, likelihood, model, Index, nparam, options.get("PCOV"))
    init = MCMC_initialize_general(setting, n_chains, init_rng or np.random.default_rng([int(seed), 7]), device)
    pb, ch = _make_chains(setting, init, device, devices)
    pr, po = _lists(setting["prior"], _PRIOR_ORDER, "prior"), _lists(setting["proposal"], _PROP_ORDER, "proposal")
    ess_temp = list(ESS_vec)
    ess_temp[4] = 0
    if standard_entries:
        o_burn = chm.make_opts(pr, po, ess_temp, update_traj=False)
        o_full = chm.make_opts(pr, po, list(ESS_vec), update_traj=True)
    else:
        ent = tuple(zip(par_ids, is_logs, pr_ids))
        # (i < burn1 runs the entries with hyper = F, i >= burn1 with options$hyper: R/LNA_chpt_slice.R:791-792, 809-810)
        o_burn = chm.make_opts_general(pr, po, ess_temp, ent, hyper=False, update_traj=False)
        o_full = chm.make_opts_general(pr, po, list(ESS_vec), ent, hyper=hyper_flag, update_traj=True)
    if ESS_vec2 is not None:  # a second parameter slice update replaces the trajectory update for i >= burn1 (:826-833)
        o_full.use_ess2 = 1
        for q in range(7):
            o_full.ESS_vec2[q] = int(ESS_vec2[q])
    burn1 = int(options.get("burn1", 5000))
    B, niter, thin = int(n_chains), int(niter), int(thin)
    rec = {"par": np.empty((B, niter, ch.npar)), "l": np.empty((B, niter)), "l1": np.empty((B, niter)),
           "traj": np.empty((B, niter // thin, pb.nrow, pb.p + 1))}
    _begin_history(ch, niter)
    for i in range(1, niter + 1):  # R's 1-based iteration counter
        ch.step_rng(o_burn if i < burn1 else o_full, seed, i, chain_offset)
        _record(ch, rec, i - 1, thin, i % thin == 0)
        if verbose and i % 100 == 0:
            print(i, float(np.mean(ch.get()["coalLog"])))
    _fetch_history(ch, rec)
    return _finish(ch, setting, rec, B, squeeze)


def General_MCMC2(coal_obs, times, t_correct, N, gridsize=1000, niter=1000, burn=0, thin=5, changetime=(),
                  DEMS=("S", "I"), prior=None, proposal=None, control=None, updateVec=(1, 1, 1, 1, 1, 1, 1),
                  likelihood="volz", model="SIR", Index=(0, 1), nparam=2, method="seq", options=None, verbose=False, *,
                  n_chains=1, seed=0, device=0, devices=None, chain_offset=0, squeeze=True, init_rng=None):
    """R/LNA_chpt_slice.R:508-693 for `n_chains` chains: stage 1 (i < burn1: scalar Metropolis entries + change-point
    slice update) and stage 2 (burn1 <= i < warmup2: + hyper-parameter Metropolis step when updateVec[5] == 0.5, +
    trajectory slice update).  Like the R code it stores the trajectory of EVERY iteration in slot i (so `niter / thin`
    must cover niter: thin = 1, or the R call fails too); here slots are filled every `thin` iterations instead."""
    if likelihood != "volz" or model != "SIR":
        raise LnaError("General_MCMC2: the GPU path has the volz likelihood with model = 'SIR' only")
    options = dict(options or {})
    warmup2 = int(options.get("warmup2") or 100000)
    burn1 = int(options.get("burn1", 5000))
    nch = len(np.ravel(changetime))
    # Update_Param_each is called with parIdlist[-4], isLoglist[-4], priorIdlist[-4] (R/LNA_chpt_slice.R:611-612): the
    # FOURTH entry is dropped, whatever the length of the lists
    drop4 = lambda seq: [v for q, v in enumerate(seq) if q != 3]
    par_ids = drop4([int(v) for v in dict(options.get("parIdlist") or {"a": 2, "b": 3, "c": 4, "d": 5 + nch}).values()])
    is_log = drop4([int(v) for v in dict(options.get("isLoglist") or {"a": 1, "b": 0, "c": 1, "d": 0}).values()])
    pr_ids = drop4([int(v) for v in dict(options.get("priorIdlist") or {"a": 1, "b": 2, "c": 3, "d": 5}).values()])
    if not (len(par_ids) == len(is_log) == len(pr_ids)) or len(par_ids) > 4:
        raise LnaError("General_MCMC2: parameters do not match")
    entries = tuple((pid - 1, lg, prid) for pid, lg, prid in zip(par_ids, is_log, pr_ids))  # R is 1-based
    setting = MCMC_setup_general(coal_obs, times, t_correct, N, gridsize, niter, burn, thin, changetime, DEMS, prior,
                                 proposal, control, likelihood, model, Index, nparam, options.get("PCOV"))
    init = MCMC_initialize_general(setting, n_chains, init_rng or np.random.default_rng([int(seed), 7]), device)
    pb, ch = _make_chains(setting, init, device, devices)
    pr, po = _lists(setting["prior"], _PRIOR_ORDER, "prior"), _lists(setting["proposal"], _PROP_ORDER, "proposal")
    upd_chpt = int(updateVec[5]) == 1
    o1 = chm.make_mcmc2_opts(pr, po, entries, update_hyper=False, update_chpt=upd_chpt, update_traj=False)
    o2 = chm.make_mcmc2_opts(pr, po, entries, update_hyper=float(updateVec[4]) == 0.5, update_chpt=upd_chpt,
                             update_traj=int(updateVec[6]) == 1)
    B, niter, thin = int(n_chains), int(niter), int(thin)
    n_norm2 = nch + pb.k * pb.p
    rec = {"par": np.empty((B, niter, ch.npar)), "l": np.empty((B, niter)), "l1": np.empty((B, niter)),
           "traj": np.empty((B, niter // thin, pb.nrow, pb.p + 1))}
    # stage 3 (i >= warmup2, :655-663): joint adaptive random walk over parIndexALL[updateVec[c(1:3,5)] > 0] (:563-569)
    uv = [float(v) for v in updateVec]
    joint_entries = tuple((jidx, lg, prid) for jidx, lg, prid, on in
                          zip((1, 2, 3, 4), (1, 0, 1, 0), (1, 2, 3, 5), (uv[0], uv[1], uv[2], uv[4])) if on > 0)
    # updateVec[4] (mu) has no parameter in the SIR model: the R code stops with 'priors do not match' INSIDE
    # update_Param_joint, i.e. only once an iteration reaches the joint stage (the default warmup2 = 100000 never does)
    joint_invalid = uv[3] > 0
    d = len(joint_entries)
    par_cols = [ch.npar - 1 if j == 4 else j for j, _, _ in joint_entries]
    o3 = chm.make_joint_opts(pr, joint_entries, update_chpt=upd_chpt, update_traj=int(updateVec[6]) == 1)
    fix_pcov = setting.get("PCOV") is not None
    tune, beta, up = float(options.get("tune") or 0.01), float(options.get("beta", 0.05)), int(options.get("up", 2000))
    pcov = None
    if fix_pcov:
        pcov = np.broadcast_to(np.asarray(setting["PCOV"], dtype=np.float64) * tune, (B, d, d)).copy()
        ch.set_joint_transform(chm.mvrnorm_transform(pcov))
    L = _lib.lib()
    _begin_history(ch, niter)
    for i in range(1, niter + 1):
        joint = i >= max(warmup2, burn1)
        if joint and joint_invalid:
            raise LnaError("General_MCMC2: priors do not match (updateVec[4] (mu) has no parameter in the SIR model; "
                           f"reached at iteration {i} >= warmup2)")
        if not fix_pcov and (i == warmup2 or (i > warmup2 and i % up == 0)):
            _fetch_history(ch, rec)  # (rows < i - 1 are recorded; the adaptation reads the chain's own past)
            # proposal covariance from the chain's own history, rows floor(i/3) .. i-1 of `params` (:608-630)
            lo = max(i // 3, 1)
            if i - lo < 2:
                raise LnaError("General_MCMC2: warmup2 is too small to estimate the proposal covariance (needs >= 2 rows)")
            hist = rec["par"][:, lo - 1:i - 1, :][:, :, par_cols].copy()
            for q, (_, lg, _) in enumerate(joint_entries):
                if lg:
                    hist[:, :, q] = np.log(hist[:, :, q])
            dev = hist - hist.mean(axis=1, keepdims=True)
            cov = np.einsum("bti,btj->bij", dev, dev) / (hist.shape[1] - 1)
            sigma_n = (2.38 ** 2) * cov / d
            pcov = (beta * sigma_n + (1 - beta) * np.eye(d) * 0.0001 / d) * tune
            ch.set_joint_transform(chm.mvrnorm_transform(pcov))
        opts_i = o3 if joint else (o1 if i < burn1 else o2)
        ch.step_mcmc2_rng(opts_i, seed, i, chain_offset)
        _record(ch, rec, i - 1, thin, i % thin == 0)
        if verbose and i % 100 == 0:
            print(i, float(np.mean(ch.get()["coalLog"])))
    _fetch_history(ch, rec)
    return _finish(ch, setting, rec, B, squeeze)
