"""Timeline of the resident-chain iteration: LNA_TIMELINE=1 LNA_GRAPH=0 python tools/timeline.py B G free stagger_us out.csv [plan]"""
import os, sys
os.environ["LNA_TIMELINE"] = "1"; os.environ["LNA_GRAPH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from lnaphylodyn_b200 import api, chains as chm, _lib
B, G, free, stag, out = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
os.environ["LNA_CHAIN_GROUPS"] = G; os.environ["LNA_STAGGER_US"] = stag
if len(sys.argv) > 6: os.environ["LNA_SLICE_PLAN"] = sys.argv[6]
g = Golden("changepoint_sim")
par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"), g.setting("control.hyper")])
pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
ch = chm.Chains(pb, np.tile(par0, (B, 1)), g.setting("control.traj"), free_running=bool(free))
ob, of = chm.vignette_opts(g, update_traj=False), chm.vignette_opts(g, update_traj=True)
for it in range(4): ch.step_rng(ob, 1, it)
for it in range(4, 12): ch.step_rng(of, 1, it)
ch.sync()
try: _lib.lib().lna_debug_timeline_dump(b"/dev/null")
except Exception: pass
if free: ch.set_free_running(False); ch.set_free_running(True)
for it in range(12, 18): ch.step_rng(of, 1, it)
ch.sync()
_lib.check(_lib.lib().lna_debug_timeline_dump(out.encode()), "dump")
print("wrote", out)
