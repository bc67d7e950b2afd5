"""BASELINE north_star: "ESlice chains step-for-step equal over a fixed horizon".  200 iterations x 64 chains per driver
body, the GPU chains on device-drawn streams (CUDA-graph replay, the automatic phase plans, stream groups) against the CPU
oracle's restated R iteration replaying the SAME streams.  Every chain-iteration is compared; decisions that differ
(accepted candidate, number of evaluations, Metropolis accepts) are COUNTED and reported, not hidden behind a tolerance:
the test fails on the first flipped decision and prints how many chain-iterations agreed before it."""
import concurrent.futures as cf
import os

import numpy as np
import pytest

from conftest import relerr
from oracle import driver

pytestmark = pytest.mark.gpu
ITERS, CHAINS = 200, 64


def _replay(bodies, streams, workers=None):
    """bodies[c](nrm, unf, it) advances chain c's oracle state; chains run in parallel threads (the C oracle releases the GIL)."""
    def run(c):
        out = []
        for it, (nrm, unf) in enumerate(streams):
            out.append(bodies[c](nrm[c], unf[c], it))
        return out
    with cf.ThreadPoolExecutor(max_workers=workers or min(32, os.cpu_count() or 4)) as ex:
        return list(ex.map(run, range(len(bodies))))


def _compare(tag, gpu_hist, ora_hist):
    """gpu_hist[it] = dict of (B, ...) arrays, ora_hist[c][it] = dict.  Returns (worst relative error, report string)."""
    worst, flips = 0.0, []
    B = len(ora_hist)
    for it, g in enumerate(gpu_hist):
        for c in range(B):
            o = ora_hist[c][it]
            dec_same = (g["n_full_evals"][c] == o["n_full"] and g["n_cheap_evals"][c] == o["n_cheap"]
                        and g["n_mh_accept"][c] == o["n_mh_accept"])
            if not dec_same:
                flips.append((it, c, int(g["n_full_evals"][c]), o["n_full"], int(g["n_mh_accept"][c]), o["n_mh_accept"]))
                continue
            worst = max(worst, relerr(g["par"][c], o["par"]), abs(g["coalLog"][c] - o["coalLog"]) / abs(o["coalLog"]))
        if flips:
            break
    agreed = (it if flips else len(gpu_hist)) * B
    msg = (f"{tag}: {agreed} chain-iterations compared step for step, flipped decisions: {len(flips)}"
           + (f", first at (iteration, chain, n_full gpu/oracle, accepts gpu/oracle) = {flips[0]}" if flips else "")
           + f"; worst relative deviation of par / coalLog {worst:.2e}")
    return worst, flips, msg


def _snap(s):
    return {k: s[k].copy() for k in ("par", "coalLog", "n_full_evals", "n_cheap_evals", "n_mh_accept")}


def _osnap(st):
    return {"par": st["par"].copy(), "coalLog": st["coalLog"], "n_full": st["n_full"], "n_cheap": st["n_cheap"],
            "n_mh_accept": st["n_mh_accept"]}


@pytest.mark.parametrize("schedule", ["auto", "barrier-free"])
def test_general_mcmc_with_eslice_200_iterations(changepoint, monkeypatch, schedule):
    """Changpoint_sim vignette body (gamma walk + no-op hyper entry, parameter slice update, trajectory slice update from
    iteration 20 on).  `auto` is what 64 chains select (speculative waves, three-warp pipeline); `barrier-free` forces the
    kernel that batches of 1007 to ~28 000 chains select (lna_async.cuh: dynamic candidate assignment, early rejection), with 16
    lanes per chain as in the lower half of that range."""
    from lnaphylodyn_b200 import api, chains as chm
    if schedule == "barrier-free":
        monkeypatch.setenv("LNA_ESS_ASYNC_W", "16")
    g = changepoint
    B, seed, burn1 = CHAINS, 2024, 20
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=True)
    streams, hist = [], []
    for it in range(ITERS):
        streams.append(ch.device_streams(seed, it))
        ch.step_rng(chm.vignette_opts(g, update_traj=it >= burn1), seed, it)
        hist.append(_snap(ch.get()))
    states = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]

    def body(c):
        def f(nrm, unf, it):
            states[c] = driver.eslice_iteration(states[c], g, driver.opts_from_golden(g, update_traj=it >= burn1), nrm, unf)
            return _osnap(states[c])
        return f
    worst, flips, msg = _compare("General_MCMC_with_ESlice", hist, _replay([body(c) for c in range(B)], streams))
    print(msg)
    assert not flips, msg
    assert worst < 1e-8, msg   # 200 iterations of feedback: rounding differences of 1e-15 per evaluation have room to grow


def test_general_mcmc2_200_iterations(constant):
    """constant_sim vignette body (three Metropolis entries as a speculative tree, hyper step, change-point slice update,
    trajectory slice update from iteration 20 on)."""
    from lnaphylodyn_b200 import _lib, api, chains as chm
    g = constant
    B, seed, burn1 = CHAINS, 77, 20
    par0 = np.concatenate([[g.x_r[0], float(g.setting("control.alpha")[0]) * g.x_r[0]], g.setting("control.R0"),
                           g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=True)
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    n2 = 39 + 80
    streams, hist = [], []
    for it in range(ITERS):
        nrm, unf = np.empty((B, n2)), np.empty((B, chm.ITER2_UNIF))
        _lib.check(_lib.lib().lna_draw_streams(pb.h, B, 0, n2, chm.ITER2_UNIF, seed, it, api._ptr(nrm), api._ptr(unf)))
        streams.append((nrm, unf))
        burn = it < burn1
        ch.step_mcmc2_rng(chm.make_mcmc2_opts(prior, proposal, update_hyper=not burn, update_traj=not burn), seed, it)
        hist.append(_snap(ch.get()))
    states = [driver.init_state(g, par0, g.setting("control.traj")) for _ in range(B)]

    def body(c):
        def f(nrm, unf, it):
            states[c] = driver.mcmc2_iteration(states[c], g, driver.mcmc2_opts_from_golden(g, burn=it < burn1), nrm, unf)
            return _osnap(states[c])
        return f
    worst, flips, msg = _compare("General_MCMC2", hist, _replay([body(c) for c in range(B)], streams))
    print(msg)
    assert not flips, msg
    assert worst < 1e-8, msg
