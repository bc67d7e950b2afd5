"""Experiment (GPU box): the balanced FULL kernel (k_full_loglik_bal, LNA_BAL=1) against the plain one (LNA_BAL=0):
bit-equality of every output on ragged batch sizes, then the launch time per batch size with CUDA events and an L2 flush
between launches.  usage: python tools/bal_probe.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from lnaphylodyn_b200 import _lib, api

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.Workload("changepoint")
pb = w.problem(api, 0)
dev = torch.device("cuda:0")
L = _lib.lib()
import ctypes as C
_lib.check(L.lna_problem_set_stream(pb.h, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "set_stream")
nC, m = 4800, 16
param, initial, eps = bench.synth_inputs(w, nC, m, seed=4242)
Bmax = nC * m
eps_t = np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))
pd = torch.from_numpy(param.reshape(Bmax, -1)).to(dev)
idv = torch.from_numpy(initial).to(dev).repeat_interleave(m, dim=0).contiguous()
ed = torch.from_numpy(eps_t.reshape(nC, -1)).to(dev).repeat_interleave(m, dim=0).contiguous()
p, k, nrow = 2, pb.k, pb.nrow


def run(B, bal, store):
    os.environ["LNA_BAL"] = bal
    cl = torch.full((B,), -7.0, dtype=torch.float64, device=dev)
    st = torch.full((B,), -7, dtype=torch.int32, device=dev)
    outs = [torch.full((B, n), -7.0, dtype=torch.float64, device=dev) for n in (p * p * k, p * p * k, nrow * (p + 1), nrow, nrow * (p + 1))] if store else [None] * 5
    ptrs = [o.data_ptr() if o is not None else None for o in outs]
    _lib.check(L.lna_full_loglik_dev(pb.h, B, pd.data_ptr(), idv.data_ptr(), ed.data_ptr(), cl.data_ptr(), st.data_ptr(), *ptrs), "lna_full_loglik_dev")
    torch.cuda.synchronize()
    return [cl, st] + [o for o in outs if o is not None]


bad = 0
for B in (1, 31, 33, 100, 1000, 4737, 4738, 9472, 20000, 65536, 70000):
    for store in (False, True):
        if store and B > 20000:
            continue
        a, b = run(B, "0", store), run(B, "1", store)
        same = all(torch.equal(x, y) for x, y in zip(a, b))
        bad += 0 if same else 1
        print(f"B={B:6d} store={int(store)} outputs bit-identical: {same}   coalLog mean {float(a[0][a[0] > -1e6].mean()):.6f}", flush=True)
print("MISMATCHES", bad)

flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
cl = torch.empty(Bmax, dtype=torch.float64, device=dev)
st = torch.empty(Bmax, dtype=torch.int32, device=dev)
print("B, loads/sub-partition, plain ms, balanced ms, ratio")
for B in (18944, 28416, 37888, 47360, 56832, 60000, 65536, 66304, 70000, 75776):
    res = {}
    for bal in ("0", "1"):
        os.environ["LNA_BAL"] = bal
        ts = []
        for r in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.lna_full_loglik_dev(pb.h, B, pd.data_ptr(), idv.data_ptr(), ed.data_ptr(), cl.data_ptr(), st.data_ptr(), None, None, None, None, None), "x")
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        res[bal] = float(np.median(ts))
    print(f"{B:6d}  {B / 32 / 592:5.2f}  {res['0']:.4f}  {res['1']:.4f}  {res['0'] / res['1']:.3f}   balanced: {B / res['1'] / 1e3:.4g} evals/s  W-frac {B * 227200 / (res['1'] * 1e-3) / 37.13e12:.3f}", flush=True)

# stress: the shared-memory hand-over (state, fence, flag / flag, fence, state) under repetition -- a lost or early hand-over
# would change bits in the cut loads
if len(sys.argv) > 2 and sys.argv[2] == "stress":
    n_bad, n_run = 0, 0
    for B in (65536, 60000, 47360, 28416, 9489, 700):
        ref = run(B, "0", False)
        for rep in range(60):
            got = run(B, "1", False)
            n_run += 1
            if not all(torch.equal(x, y) for x, y in zip(ref, got)):
                n_bad += 1
    print(f"stress: {n_run} balanced launches compared with the plain kernel's results, {n_bad} mismatching", flush=True)
