"""Per-kernel durations of one resident-chain iteration (LNA_TIMELINE), wave kernels against the barrier-free slice kernel,
and the executed-work counters.  usage: python tools/async_timeline.py [B]"""
import os, subprocess, sys, csv, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = sys.argv[1] if len(sys.argv) > 1 else "4096"
if len(sys.argv) > 2 and sys.argv[2] == "counts":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from conftest import Golden
    from lnaphylodyn_b200 import api, chains as chm
    g = Golden("changepoint_sim")
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
    pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(par0, (int(B), 1)), g.setting("control.traj"))
    of = chm.vignette_opts(g, update_traj=True)
    for it in range(12): ch.step_rng(of, 1, it)
    e0 = ch.exec_counts(); s0 = ch.get()
    for it in range(12, 22): ch.step_rng(of, 1, it)
    e1 = ch.exec_counts(); s1 = ch.get()
    n = 10 * int(B)
    ne = (s1["n_full_evals"] - s0["n_full_evals"])
    print(f"  consumed FULL evals per chain-iteration {ne.sum() / n:.2f}; started {(e1[0] - e0[0]) / n:.2f}; warp-evaluations per iteration "
          f"{(e1[1] - e0[1]) / 10:.0f}; slice candidates needed: mean {ne.mean() / 10 - 2:.2f}, P(>8) {np.mean(ne / 10 - 2 > 8):.2f}, P(>16) {np.mean(ne / 10 - 2 > 16):.3f}")
    sys.exit(0)
for label, env in (("wave kernels", {"LNA_ESS_ASYNC": "0"}), ("async W=8", {"LNA_ESS_ASYNC_W": "8"}), ("async W=4", {"LNA_ESS_ASYNC_W": "4"})):
    out = f"/tmp/tl_{label.replace(' ', '_').replace('=', '')}.csv"
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "timeline.py"), B, "1", "0", "0", out], env=e, capture_output=True, text=True)
    if r.returncode:
        print(label, "FAILED", r.stderr[-400:]); continue
    print(label)
    rows = list(csv.DictReader(open(out)))
    dur = collections.OrderedDict()
    prev = {}
    for row in rows:  # consecutive end-of-kernel timestamps of one stream give each kernel's span
        st, k, t = row["stream"], row["kernel"], float(row["t_end_us"])
        if st in prev:
            dur.setdefault(k, []).append(t - prev[st])
        prev[st] = t
    tot = 0.0
    for k, v in dur.items():
        print(f"   {k:40s} n={len(v):3d} mean {sum(v) / len(v):8.1f} us")
        tot += sum(v)
    print(f"   total per iteration {tot / 6:.1f} us")
    subprocess.run([sys.executable, __file__, B, "counts"], env=e)
