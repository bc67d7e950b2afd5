"""Experiment: split B chains into G groups on separate streams so latency-bound stages overlap."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = 10
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
dev = torch.device("cuda", 0)
for W in (0, -16):
  for G in (1, 2, 4, 8):
    Bg = B // G
    pbs = [api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init) for _ in range(G)]
    chs = [chm.Chains(pb, np.tile(par0, (Bg, 1)), g.setting("control.traj")) for pb in pbs]
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    nrm = torch.randn((iters + 4, G, Bg, chs[0].n_norm), dtype=torch.float64, device=dev, generator=gen)
    unf = torch.rand((iters + 4, G, Bg, chm.ITER_UNIF), dtype=torch.float64, device=dev, generator=gen).clamp_(1e-300, 1 - 1e-16)
    opts = chm.vignette_opts(g, update_traj=True, spec_width=W)
    torch.cuda.synchronize()
    for it in range(4):
        for k, ch in enumerate(chs):
            ch.step_dev(opts, nrm[it, k].data_ptr(), unf[it, k].data_ptr())
    for pb in pbs: pb.sync()
    t0 = time.perf_counter()
    for it in range(4, 4 + iters):
        for k, ch in enumerate(chs):
            ch.step_dev(opts, nrm[it, k].data_ptr(), unf[it, k].data_ptr())
    for pb in pbs: pb.sync()
    dt = time.perf_counter() - t0
    ne = sum(ch.get()["n_full_evals"].sum() for ch in chs) / (B * (iters + 4))
    print(f"W={W} G={G}: {1e3*dt/iters:.3f} ms/iter  {B*iters/dt/1e6:.2f} M chain-iter/s  (full evals/iter {ne:.2f})")
    del chs, pbs
