"""Run the Geweke-style joint-distribution check (lnaphylodyn_b200/geweke.py) and print its verdict.
usage: python tools/geweke_probe.py [n_chains] [n_iter] [I0]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lnaphylodyn_b200 import geweke

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 60
I0 = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
t0 = time.time()
r = geweke.run(nc, ni, seed=3, I0=I0)
dt = time.time() - t0
z, d = geweke.z_scores(r)
print(f"{nc} chains x {ni} iterations, I0 = {I0}: {r['genealogies']} genealogies simulated, {dt:.1f} s "
      f"({1e3 * dt / r['genealogies']:.1f} ms per iteration)")
print(f"forward draws valid: {r['forward_valid_frac']:.4f}; root collapse in {r['root_collapse_frac']:.3f} of the genealogies; "
      f"negative latent state after a transition: {r['negative_state_frac']:.4f}")
for n, zz, dd, sd, fm, cm in zip(r["names"], z, d, r["prior_sd"], r["forward"].mean(0), r["chain_means"].mean(0)):
    print(f"  {n:12s} forward mean {fm:+.4f}  chains {cm:+.4f}  diff {dd:+.4f} ({dd / sd if sd == sd else float('nan'):+.2f} prior sd)  z = {zz:+.2f}")
