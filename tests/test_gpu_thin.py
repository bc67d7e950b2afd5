"""volz_thin_sir (R/coal_simulation.R:11-70) on the device -- one warp per genealogy, 32 thinning proposals per
round -- against the sequential CPU restatement consuming THE SAME counter-based stream (bit-identical coalescent times and
lineage counts), against the NumPy restatement in distribution, and at the configs[4] size (5000 tips)."""
import time

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import sampler

pytestmark = pytest.mark.gpu


def sir_traj(N, n=2001):
    times = np.linspace(0, 100, n)  # (t_correct = 90 / 60 below: the grid extends beyond it, as volz_thin_sir needs)
    ode = orc.ODE_rk45([N, 1.0], times, [2.2, 0.2], [N], [0, 2, 0, 1])
    return ode, orc.betaTs([2.2, 0.2], times, [N], [0, 2, 0, 1])


def test_device_thinning_is_the_sequential_loop_bit_for_bit():
    from lnaphylodyn_b200 import genealogy as gen
    ode, betaN = sir_traj(1e5)
    design = ([0, 10, 20, 40, 80, 85], [100, 100, 145, 145, 9, 1])
    for seed, item in ((11, 0), (11, 3), (2024, 0)):
        d = gen.volz_thin_sir(ode, 90.0, *design, betaN, seed=seed, item_offset=item)
        assert d["status"] == 0 and len(d["coal_times"]) == 499
        ua, ub = gen.thin_draws(seed, item, 0, d["n_draws"])
        o = orc.volz_thin_sir(ode, 90.0, *design, betaN, ua=ua, ub=ub)
        assert o["status"] == 0 and o["n_draws"] == d["n_draws"]
        assert np.array_equal(o["coal_times"], d["coal_times"]) and np.array_equal(o["lineages"], d["lineages"])
    # `lower` below the smallest thed: fewer, better proposals -- same law, still the sequential loop bit for bit
    thed = 1.0 / (betaN * ode[:, 1] / (ode[:, 2] / 2.0))
    low = 0.5 * thed[: 1801].min()
    d = gen.volz_thin_sir(ode, 90.0, *design, betaN, seed=5, lower=low)
    ua, ub = gen.thin_draws(5, 0, 0, d["n_draws"])
    o = orc.volz_thin_sir(ode, 90.0, *design, betaN, lower=low, ua=ua, ub=ub)
    assert np.array_equal(o["coal_times"], d["coal_times"]) and o["n_draws"] == d["n_draws"]
    # a batch = the single calls with the same stream ids; per-item trajectories
    many = gen.volz_thin_sir(ode, 90.0, *design, betaN, seed=11, n_genealogies=6)
    one = gen.volz_thin_sir(ode, 90.0, *design, betaN, seed=11, item_offset=3)
    assert np.array_equal(many[3]["coal_times"], one["coal_times"])
    assert not np.array_equal(many[2]["coal_times"], many[3]["coal_times"])
    trajs = np.stack([ode, ode * np.array([1.0, 1.0, 0.5])])
    two = gen.volz_thin_sir(trajs, 90.0, *design, np.stack([betaN, betaN]), seed=11)
    assert np.array_equal(two[0]["coal_times"], many[0]["coal_times"])
    assert two[1]["coal_times"].mean() < two[0]["coal_times"].mean()  # half the prevalence: lineages coalesce sooner


def test_device_thinning_has_the_law_of_the_numpy_restatement():
    """512 device genealogies against 512 NumPy ones (20 tips each): mean and spread of the tree height and of the first
    coalescence time agree within 5 standard errors."""
    from lnaphylodyn_b200 import genealogy as gen
    ode, betaN = sir_traj(1e4)
    design = ([0, 15], [12, 8])
    R = 512
    dev = gen.volz_thin_sir(ode, 60.0, *design, betaN, seed=77, n_genealogies=R)
    rng = np.random.default_rng(123)
    ref = [sampler.volz_thin_sir(ode, 60.0, *design, betaN, rng) for _ in range(R)]
    for stat in (lambda c: c["coal_times"][-1], lambda c: c["coal_times"][0], lambda c: np.mean(c["coal_times"])):
        a, b = np.array([stat(c) for c in dev]), np.array([stat(c) for c in ref])
        se = np.sqrt(a.var() / R + b.var() / R)
        assert abs(a.mean() - b.mean()) < 5 * se, (a.mean(), b.mean(), se)
        assert 0.7 < a.std() / b.std() < 1.4
    assert all(len(c["coal_times"]) == 19 for c in dev)


def test_configs4_genealogy_on_the_device():
    """The configs[4] synthetic genealogy (5000 tips on the deterministic N = 1e6 epidemic, lower = 1 as in the reference):
    4999 sorted coalescent times, valid input of coal_lik_init; seconds on the device where the NumPy loop takes minutes."""
    from lnaphylodyn_b200 import genealogy as gen
    ode, betaN = sir_traj(1e6)
    design = ([0, 10, 20, 40, 80, 85], [1000, 1000, 1450, 1450, 90, 10])
    t0 = time.time()
    d = gen.volz_thin_sir(ode, 90.0, *design, betaN, seed=20260)
    dt = time.time() - t0
    assert d["status"] == 0 and len(d["coal_times"]) == 4999
    assert np.all(np.diff(d["coal_times"]) >= 0) and d["coal_times"].max() <= 90.0
    assert np.array_equal(np.sort(d["lineages"])[::-1][:3], np.sort(d["lineages"])[::-1][:3]) and d["lineages"].min() >= 2
    ini = gen.coal_lik_init(d["samp_times"], d["n_sampled"], d["coal_times"], ode[::50, 0])
    assert ini["gridrep"].sum() == len(ini["C"]) and ini["y"].sum() == 4999
    # same law as the committed NumPy fixture: the median coalescent time of 4999 events
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "synthetic_5000.npz"))
    assert abs(np.median(d["coal_times"]) - np.median(z["coal_times"])) < 1.0
    print(f"volz_thin_sir, 5000 tips: {d['n_draws']:.3e} proposals in {dt:.2f} s on one warp")
