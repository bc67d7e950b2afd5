"""Geweke-style joint-distribution check on the device (lnaphylodyn_b200/geweke.py; the reference's versions for
its superseded drivers are R/testfuntions.R:1-197).  The successive-conditional simulator -- device MCMC transitions given the
genealogy, alternated with device volz_thin_sir genealogies given parameters and trajectory -- must reproduce the law the
marginal-conditional simulator (prior draws) gives to functionals of the parameters.  A negative control (genealogies
simulated with a cell convention the likelihood does not use) shows the check has the power to see such a mismatch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_geweke_joint_distribution_check():
    from lnaphylodyn_b200 import geweke
    r = geweke.run(n_chains=16, n_iter=40, seed=3)
    z, d = geweke.z_scores(r)
    rel = d[:3] / r["prior_sd"][:3]
    print("Geweke check:", ", ".join(f"{n}: z = {zz:+.2f} (diff {dd:+.4f})" for n, zz, dd in zip(r["names"], z, d)),
          f"; {r['genealogies']} genealogies, root collapse in {r['root_collapse_frac']:.2f} of them, forward draws valid "
          f"{r['forward_valid_frac']:.4f}")
    assert r["forward_valid_frac"] > 0.98 and r["negative_state_frac"] < 0.01
    assert np.all(np.abs(z) < 4.0), (z, d)
    assert np.all(np.abs(rel) < 0.35), rel  # (the root-collapse tilt is ~ +0.2 sd^2 of log R0: see geweke.py)
    # negative control: the simulator's own cell convention on the coarse grid halves the effective coalescent rate the
    # likelihood sees -> the chains drift away from the prior, and the check says so
    bad = geweke.run(n_chains=8, n_iter=30, seed=3, likelihood_cells=False)
    zb, db = geweke.z_scores(bad)
    print("negative control:", ", ".join(f"{n}: z = {zz:+.2f}" for n, zz in zip(bad["names"], zb)))
    assert np.max(np.abs(zb[:2])) > 4.0, zb
