"""TEST INFRASTRUCTURE ONLY (same rule as oracle/oracle.py): the reference's R-level data simulator restated in NumPy.

volz_thin_sir (R/coal_simulation.R:11-70) with a NumPy Generator standing in for R's rexp / runif.  Used by
tests/golden/make_synthetic.py (the committed configs[4] fixture synthetic_5000.npz was drawn with default_rng(20260) through
this loop) and as the distributional reference of the device simulator (lnaphylodyn_b200.genealogy.volz_thin_sir).  The
stream-replaying C restatement that pins the device kernel bit for bit is oracle.volz_thin_sir (oracle/lna_oracle.c)."""
import numpy as np


def volz_thin_sir(eff_pop, t_correct, samp_times, n_sampled, betaN, rng, lower=1.0):
    """R/coal_simulation.R:11-70: simulate coalescent times by thinning against the Volz rate on an SIR
    trajectory `eff_pop` (columns: time, S, I).  `rng` is a numpy Generator (replaces R's rexp/runif)."""
    eff_pop = np.asarray(eff_pop, dtype=np.float64)
    samp_times = np.asarray(samp_times, dtype=np.float64)
    n_sampled = np.asarray(n_sampled, dtype=np.int64)
    coal_times, lineages = [], []
    curr = 0
    active = int(n_sampled[curr])
    time = float(samp_times[curr])
    traj = 1.0 / (betaN * eff_pop[:, 1] / (eff_pop[:, 2] / 2.0))
    ts = t_correct - eff_pop[:, 0]
    if not eff_pop[-1, 0] > t_correct:
        raise ValueError("volz_thin_sir: the time grid must extend beyond t_correct (R stops with an error on ts[0])")
    i = int(np.argmin(ts >= 0)) - 1                # which.min(ts >= 0) - 1, then 0-based = that - 1 + ... see below
    # R: i = which.min(ts>=0) - 1 is a 1-based index of the last grid time <= t_correct; 0-based: subtract 1
    i = max(i, 0)
    max_samp = samp_times.max()
    while time <= max_samp or active > 1:
        if active == 1:
            curr += 1
            active += int(n_sampled[curr])
            time = float(samp_times[curr])
        time = time + rng.exponential(1.0 / (0.5 * active * (active - 1) / lower))
        while time > ts[i]:
            i -= 1
            if i < 0:
                i = 0
                break
        thed = traj[i]
        if curr < len(samp_times) - 1 and time >= samp_times[curr + 1]:
            curr += 1
            active += int(n_sampled[curr])
            time = float(samp_times[curr])
        elif rng.random() <= lower / thed:
            time = min(time, t_correct)
            coal_times.append(time)
            lineages.append(active)
            active -= 1
    coal_times = np.array(coal_times)
    return {"coal_times": coal_times, "lineages": np.array(lineages), "samp_times": samp_times,
            "n_sampled": n_sampled.astype(np.float64)}
