"""Launch the bench's headline kernel -- k_full_loglik<0, 0>, B = 4096 chains x 16 proposals = 65 536 FULL evaluations of the
Changpoint_sim problem, device-resident per-item inputs -- a few times and nothing else (the target of the ncu --set full
capture whose DRAM bytes bench.py reports as roofline.traffic).  usage: python tools/headline_kernel.py [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from lnaphylodyn_b200 import _lib, api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w = bench.Workload("changepoint")
pb = w.problem(api, 0)
nC, m = 4096, 16
B = nC * m
dev = torch.device("cuda:0")
bufs = []
for i in range(2):
    param, initial, eps = bench.synth_inputs(w, nC, m, seed=10_000 + i)
    eps_t = np.ascontiguousarray(np.transpose(eps, (0, 2, 1)))
    pd = torch.from_numpy(param.reshape(B, -1)).to(dev)
    idv = torch.from_numpy(initial).to(dev).repeat_interleave(m, dim=0).contiguous()
    ed = torch.from_numpy(eps_t.reshape(nC, -1)).to(dev).repeat_interleave(m, dim=0).contiguous()
    bufs.append((pd, idv, ed))
cl = torch.empty(B, dtype=torch.float64, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
L = _lib.lib()
for i in range(n):
    flush.zero_()
    p_, i_, e_ = bufs[i % 2]
    _lib.check(L.lna_full_loglik_dev(pb.h, B, p_.data_ptr(), i_.data_ptr(), e_.data_ptr(), cl.data_ptr(), st.data_ptr(),
                                     None, None, None, None, None), "lna_full_loglik_dev")
torch.cuda.synchronize()
print("ok", float(cl[cl > -1e6].mean()))
