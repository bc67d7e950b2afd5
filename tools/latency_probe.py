"""Time lna_full_loglik_dev at several batch sizes (latency-bound ... throughput-bound)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
so = sys.argv[1] if len(sys.argv) > 1 else None
from lnaphylodyn_b200 import _lib, build as _b
if so:
    _b.build = lambda *a, **k: so
from lnaphylodyn_b200 import api
import bench
g = bench.load_workload()
L = _lib.lib()
dev = torch.device("cuda", 0)
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
st = torch.cuda.current_stream(dev)
L.lna_problem_set_stream(pb.h, C.c_void_p(st.cuda_stream))
out = []
for B in (4096, 16384, 65536, 131072):
    par, eps = bench.synth_inputs(g, B, 1)
    p_ = torch.from_numpy(np.ascontiguousarray(par[:, 2:])).to(dev); i_ = torch.from_numpy(np.ascontiguousarray(par[:, :2])).to(dev)
    e_ = torch.from_numpy(np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))).to(dev)
    cl = torch.empty(B, dtype=torch.float64, device=dev); s_ = torch.empty(B, dtype=torch.int32, device=dev)
    run = lambda: L.lna_full_loglik_dev(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cl.data_ptr(), s_.data_ptr(), None, None, None, None, None)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): run()
    e1.record(st); torch.cuda.synchronize()
    out.append(f"B={B}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
print(os.path.basename(so or "default"), " | ".join(out), " checksum", float(cl.sum()))
