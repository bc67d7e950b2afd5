"""Generate tests/golden/synthetic_5000.npz: the configs[4] synthetic genealogy.

volz_thin_sir (R/coal_simulation.R:11-70, restated in NumPy in oracle/sampler.py) on the deterministic
SIR trajectory N = 1e6, R0 = 2.2, gamma = 0.2, I0 = 1 over [0, 100] (4001-point grid), t_correct = 90,
samp_times (0,10,20,40,80,85), n_sampled (1000,1000,1450,1450,90,10) = 5000 tips, NumPy default_rng(20260).
The thinning loop rejects millions of proposals (pure Python: ~8 minutes), hence the committed fixture.
Run from the repo root:  python tests/golden/make_synthetic.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import sampler as gen  # noqa: E402
from oracle import oracle as orc  # noqa: E402

times = np.linspace(0, 100, 2001)
ode = orc.ODE_rk45([1e6, 1.0], times, [2.2, 0.2], [1e6], [0, 2, 0, 1])
betaN = orc.betaTs([2.2, 0.2], times, [1e6], [0, 2, 0, 1])
c = gen.volz_thin_sir(ode, 90.0, [0, 10, 20, 40, 80, 85], [1000, 1000, 1450, 1450, 90, 10], betaN,
                      np.random.default_rng(20260))
np.savez_compressed(os.path.join(HERE, "synthetic_5000.npz"), coal_times=c["coal_times"], samp_times=c["samp_times"],
                    n_sampled=c["n_sampled"])
print(len(c["coal_times"]), "coalescent times written")
