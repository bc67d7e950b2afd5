"""Geweke-style joint-distribution check of the sampler (the reference's versions for its superseded
drivers: R/testfuntions.R:1-197, `Geweke_S2_Traj` and friends).

Two simulators of the joint law of (parameters, latent trajectory, genealogy):
  * marginal-conditional: parameters and latent increments from their priors (the batch is one FULL-evaluation launch);
  * successive-conditional: one chain per replicate that alternates the device MCMC transition `General_MCMC_with_ESlice`
    (Metropolis entries + parameter and trajectory elliptical-slice updates) GIVEN the genealogy with a fresh genealogy
    GIVEN the parameters and the trajectory, drawn on the device by `volz_thin_sir` (csrc/lna_thin.cuh).
If transition kernels, likelihood and simulator describe the same model, functionals of the parameters have the same law
under both.  Every chain is started from a marginal-conditional draw, so it is stationary from iteration 0 and the chains'
time averages are independent unbiased estimates: the comparison needs no mixing argument.

The genealogy is simulated on the COARSE latent trajectory with the piecewise-constant rate convention of the likelihood
(cell i of `volz_loglik_nh2` reads trajectory row L - i, the cell's later endpoint, LNA_functional.cpp:1140-1150, while
`volz_thin_sir` looks up the earlier endpoint of the cell a time falls in, coal_simulation.R:40-49): the rows handed to the
simulator are shifted by one cell.

What the check can NOT be: exact.  `volz_thin_sir` collapses every lineage still alive at t_correct into the root
(`time = min(time, t_correct)`, :56), and `coal_lik_init` then scores that as ONE ordinary coalescence with density
lambda(last cell) -- a factor the simulator never paid.  With I0 individuals at the origin the last pair survives to the root
with probability ~ exp(-2 R0 / (I0 (R0 - 1))), not negligible for any I0 at which the LNA is meaningful, so the two laws
differ by a known, small tilt (reported as `root_collapse_frac`).  The check therefore detects gross inconsistencies (a
misplaced factor in the coalescent rate moves log R0 by several prior standard deviations), not second-order ones."""
import numpy as np

from . import api, chains as chm, genealogy as gen


def _shifted(traj, betaN, shift=True):
    """rows for the simulator: (time_i, S_{i+1}, I_{i+1}), beta_{i+1} -- the likelihood's cell convention -- plus a closing
    row past t_correct (volz_thin_sir starts its cell search from the last grid time <= t_correct, which must not be the
    end of the grid: `which.min(ts >= 0) - 1`, coal_simulation.R:25)"""
    sh = traj.copy()
    b = betaN.copy()
    if shift:
        sh[:-1, 1:] = traj[1:, 1:]
        b[:-1] = betaN[1:]
    last = sh[-1].copy()
    last[0] += sh[1, 0] - sh[0, 0]
    return np.vstack([sh, last]), np.append(b, b[-1])


def run(n_chains=16, n_iter=60, seed=1, *, T=40.0, n_fine=800, gridsize=50, N=1e5, I0=20.0, changetime=(10.0, 20.0, 30.0),
        R0_pr=(np.log(2.2), 0.2), gamma_pr=(np.log(0.25), 0.1), hyper_pr=(3.0, 0.2), gamma_prop=0.05,
        samp_times=(0.0, 4.0, 8.0), n_sampled=(30, 30, 20), n_forward=4096, device=0, likelihood_cells=True):
    """Returns a dict with the two simulators' draws of (log R0, log gamma, log hyper, mean log ch):
    `forward` (n_forward x 4) and `chain_means` (n_chains x 4: each chain's time average), plus diagnostics.
    likelihood_cells = False simulates the genealogies with volz_thin_sir's own (earlier-endpoint) cell convention on the
    coarse grid: a deliberately inconsistent pair, the negative control that shows what the check can see."""
    times = np.linspace(0.0, T, n_fine + 1)
    cht = np.asarray(changetime, dtype=np.float64)
    nch = len(cht)
    x_r, x_i = np.concatenate([[N], cht]), [nch, 2, 0, 1]
    npar = 2 + 2 + nch + 1
    prior = [np.array([1.0, 1.0]), np.asarray(R0_pr), np.asarray(gamma_pr), np.array([0.0, 1.0]), np.asarray(hyper_pr)]
    proposal = [np.array([1.0]), np.array([0.1]), np.array([gamma_prop]), np.array([0.1]), np.array([0.1])]
    opts = chm.make_opts(prior, proposal, [0, 1, 0, 0, 1, 1, 0], update_traj=True, hyper_noop_mh=True)
    pb0 = api.Problem(times, x_r, x_i, gridsize, T, None, device=device)
    k = pb0.k
    coarse_t = times[::gridsize]

    def prior_draws(rng, n):
        par = np.empty((n, npar))
        par[:, 0], par[:, 1] = N, I0
        par[:, 2] = np.exp(rng.normal(R0_pr[0], R0_pr[1], n))
        par[:, 3] = np.exp(rng.normal(gamma_pr[0], gamma_pr[1], n))
        par[:, -1] = np.exp(rng.normal(hyper_pr[0], hyper_pr[1], n))
        par[:, 4:4 + nch] = np.exp(rng.standard_normal((n, nch)) / par[:, -1:])
        return par, rng.standard_normal((n, k, 2))

    def functionals(par):
        par = np.atleast_2d(par)
        return np.stack([np.log(par[:, 2]), np.log(par[:, 3]), np.log(par[:, -1]), np.log(par[:, 4:4 + nch]).mean(1)], 1)

    def latent(par, eps):
        """trajectories and betaN of a batch of draws; `ok` = no negative state anywhere"""
        res = pb0.full_loglik(par[:, 2:], par[:, :2], eps, want=("LatentTraj", "betaN"))
        lat = np.asarray(res["LatentTraj"])
        return lat, np.asarray(res["betaN"]), np.min(lat[:, :, 1:], axis=(1, 2)) > 0.0

    # marginal-conditional simulator
    rng = np.random.default_rng([seed, 0])
    par_f, eps_f = prior_draws(rng, n_forward)
    lat_f, bet_f, ok_f = latent(par_f, eps_f)
    forward = functionals(par_f[ok_f])

    # successive-conditional simulator
    stats = np.zeros((n_chains, 4))
    collapse, negative, item = 0, 0, 0
    for c in range(n_chains):
        crng = np.random.default_rng([seed, 1, c])
        while True:  # a valid draw of the joint prior
            par, eps = prior_draws(crng, 1)
            lat, bet, ok = latent(par, eps)
            if ok[0]:
                break
        par, eps, lat, bet = par[0], eps[0], lat[0], bet[0]
        acc = np.zeros(4)
        for it in range(n_iter):
            sh, bs = _shifted(lat, bet, likelihood_cells)
            g = gen.volz_thin_sir(sh, T, samp_times, n_sampled, bs, seed=seed, item_offset=item, device=device)
            item += 1
            collapse += int(np.sum(g["coal_times"] >= T) > 0)
            init = gen.coal_lik_init(g["samp_times"], g["n_sampled"], g["coal_times"], coarse_t)
            pb = api.Problem(times, x_r, x_i, gridsize, T, init, device=device)
            ch = chm.Chains(pb, par[None], eps)
            ch.step_rng(opts, seed, it, chain_offset=c)
            s = ch.get()
            par, eps, lat = s["par"][0], s["OriginTraj"][0], s["LatentTraj"][0]
            bet = api.betaTs(par[2:], coarse_t, x_r, x_i)
            negative += int(np.min(lat[:, 1:]) <= 0.0)
            acc += functionals(par)[0]
            del ch, pb
        stats[c] = acc / n_iter
    return {"forward": forward, "chain_means": stats, "names": ("log R0", "log gamma", "log hyper", "mean log ch"),
            "prior_sd": np.array([R0_pr[1], gamma_pr[1], hyper_pr[1], np.nan]), "forward_valid_frac": float(ok_f.mean()),
            "root_collapse_frac": collapse / (n_chains * n_iter), "negative_state_frac": negative / (n_chains * n_iter),
            "genealogies": item}


def z_scores(res):
    """(chain-mean average - forward mean) / standard error, per functional"""
    f, c = res["forward"], res["chain_means"]
    se = np.sqrt(f.var(0, ddof=1) / len(f) + c.var(0, ddof=1) / len(c))
    return (c.mean(0) - f.mean(0)) / se, c.mean(0) - f.mean(0)
