"""The batched entry of the reference-side binding (integration/rcpp_chains.cpp: General_MCMC_with_ESlice_chains) -- MANY chains
from one R session, on one or several devices -- compiled against the stand-in Rcpp headers and driven through
oracle/ref_harness_legacy.cpp (ref_chains_run).  Its records must be, bit for bit, what the Python driver layer
(lnaphylodyn_b200.chains) produces for the same seed, and 12 of its chain-iterations are replayed through the CPU oracle's
restated R iteration on the same device-drawn streams."""
import ctypes as C

import numpy as np
import pytest

from conftest import relerr
from oracle import driver
from oracle import oracle as orc
from oracle import reference as ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.dropin_available(), reason="oracle/_ref/liblna_dropin.so not built")]


def run_binding(g, par0, eps0, niter, burn1, seed, devices, ess=(0, 1, 0, 0, 1, 1, 0)):
    lib = ref.dropin().lib()._dll
    B, npar = par0.shape
    cfg, ini = orc.Cfg(g.x_r, g.x_i, npar - 2), orc.Init(g.init)
    prior = np.concatenate([np.asarray(g.setting("prior." + k), dtype=np.float64).ravel()[:2]
                            for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")])
    prop = np.array([float(np.ravel(g.setting("proposal." + k))[0]) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")])
    times = orc._d(g.times)
    k = (len(times) - 1) // g.gridsize
    nrow = k + 1
    essv = (C.c_int * 7)(*ess)
    dev = (C.c_int * len(devices))(*devices)
    ph, l, l1 = np.empty((B, npar, niter)), np.empty((B, niter)), np.empty((B, niter))
    pf, lat, cl = np.empty((B, npar)), np.empty((B, 3, nrow)), np.empty(B)
    p0, e0 = orc._d(par0), orc._d(np.asarray(eps0).T)  # k x p column-major
    lib.ref_clear_error()
    rc = lib.ref_chains_run(cfg.ref, ini.ref, orc._p(times), len(times), int(g.gridsize), C.c_double(g.t_correct), orc._p(prior),
                            orc._p(prop), essv, C.c_long(B), orc._p(p0), orc._p(e0), 1, niter, burn1, 1, C.c_double(seed), dev,
                            len(devices), orc._p(ph), orc._p(l), orc._p(l1), orc._p(pf), orc._p(lat), orc._p(cl))
    assert rc == 0, lib.ref_last_error()
    # column-major (niter, npar, B) -> [chain, iteration, entry]
    return {"par": ph.transpose(0, 2, 1), "l": l, "l1": l1, "par_final": pf, "LatentTraj": lat.transpose(0, 2, 1), "coalLog": cl}


def test_batched_binding_equals_the_python_driver_and_the_oracle(changepoint):
    from lnaphylodyn_b200 import api, chains as chm
    g = changepoint
    B, niter, burn1, seed = 96, 24, 9, 404
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    par0 = np.tile(par0, (B, 1))
    par0[:, 2] *= np.exp(0.05 * np.random.default_rng(1).standard_normal(B))
    eps0 = g.setting("control.traj")
    one = run_binding(g, par0, eps0, niter, burn1, seed, [0])
    # (1) the same run on two device slots (one GPU listed twice): chains are keyed by their global index -> same bits
    two = run_binding(g, par0, eps0, niter, burn1, seed, [0, 0])
    for k in ("par", "l", "l1", "LatentTraj", "coalLog"):
        assert np.array_equal(one[k], two[k]), k
    # (2) the Python driver layer, same seed, R's 1-based iteration counter
    pb = api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, par0, eps0)
    hist, streams = [], []
    for i in range(1, niter + 1):
        streams.append(ch.device_streams(seed, i))
        ch.step_rng(chm.vignette_opts(g, update_traj=i >= burn1), seed, i)
        s = ch.get()
        hist.append((s["par"].copy(), s["coalLog"].copy(), s["logOrigin"].copy()))
    for i in range(niter):
        assert np.array_equal(one["par"][:, i, :], hist[i][0]), i
        assert np.array_equal(one["l1"][:, i], hist[i][1]) and np.array_equal(one["l"][:, i], hist[i][2])
    assert np.array_equal(one["par_final"], hist[-1][0])
    # (3) chains 0, 40, 95 replayed through the oracle's restated R iteration on the same streams
    for c in (0, 40, 95):
        st = driver.init_state(g, par0[c], eps0)
        for i in range(niter):
            nrm, unf = streams[i]
            st = driver.eslice_iteration(st, g, driver.opts_from_golden(g, update_traj=(i + 1) >= burn1), nrm[c], unf[c])
            assert relerr(one["par"][c, i], st["par"]) < 1e-9
            assert abs(one["l1"][c, i] - st["coalLog"]) < 1e-9 * abs(st["coalLog"])
