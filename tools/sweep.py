"""configs[4] scaling sweep on one GPU: synthetic 5000-tip genealogy (tests/golden/synthetic_5000.npz), 2000 fine
steps, gridsize 50 (k = 40 LNA intervals), 39 change points; 1k..64k concurrent chains.
Prints a markdown table: FULL evaluations/s at batch = chains (one proposal per chain) and ESlice
chain-iterations/s (streams drawn on the device)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden, GOLDEN
from lnaphylodyn_b200 import _lib, api, chains as chm, genealogy as gen

g = Golden("changepoint_sim")
times = np.linspace(0, 100, 2001)
c = np.load(os.path.join(GOLDEN, "synthetic_5000.npz"))
init = gen.coal_lik_init(c["samp_times"], c["n_sampled"], c["coal_times"], times[::50])
cht = times[::50][1:-1]
x_r, x_i = np.concatenate([[1e6], cht]), [39, 2, 0, 1]
dev = torch.device("cuda", 0)
L = _lib.lib()
pb = api.Problem(times, x_r, x_i, 50, 90.0, init)
st = torch.cuda.current_stream(dev)
L.lna_problem_set_stream(pb.h, C.c_void_p(st.cuda_stream))
par0 = np.concatenate([[1e6, 1.0, 2.2, 0.2], np.ones(39), [20.0]])
opts = chm.vignette_opts(g, update_traj=True)
opts_b = chm.vignette_opts(g, update_traj=False)
print(f"genealogy: {len(c['coal_times']) + 1} tips, E = {len(init['C'])} events, ng = {init['ng']}")
print("| chains | FULL evals/s (batch = chains) | us per batch | ESlice chain-iter/s | ms per iteration | FULL evals per iteration |")
print("|---:|---:|---:|---:|---:|---:|")
for B in (1, 32, 256, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    ch = chm.Chains(pb, np.tile(par0, (B, 1)), np.zeros((40, 2)))
    for it in range(6):
        ch.step_rng(opts_b if it < 3 else opts, 5, it)
    pb.sync()
    n0 = ch.get()["n_full_evals"].sum()
    iters = 10
    t0 = time.perf_counter()
    for it in range(iters):
        ch.step_rng(opts, 5, 100 + it)
    pb.sync()
    dt = time.perf_counter() - t0
    s = ch.get()
    nfe = (s["n_full_evals"].sum() - n0) / (B * iters)
    # FULL evaluation batch of the chains' current states
    p_ = torch.from_numpy(np.ascontiguousarray(s["par"][:, 2:])).to(dev); i_ = torch.from_numpy(np.ascontiguousarray(s["par"][:, :2])).to(dev)
    e_ = torch.from_numpy(np.ascontiguousarray(np.transpose(s["OriginTraj"], (0, 2, 1)))).to(dev)
    cl = torch.empty(B, dtype=torch.float64, device=dev); s_ = torch.empty(B, dtype=torch.int32, device=dev)
    run = lambda: L.lna_full_loglik_dev(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cl.data_ptr(), s_.data_ptr(), None, None, None, None, None)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): run()
    e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    assert torch.allclose(cl.cpu(), torch.from_numpy(s["coalLog"]), rtol=1e-12)
    print(f"| {B} | {B / us * 1e6:.3e} | {us:.0f} | {B * iters / dt:.3e} | {dt / iters * 1e3:.2f} | {nfe:.1f} |")
    del ch
