"""Host-side genealogy ingestion against the reference's cached `Init` lists and the bundled Ebola tree."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lnaphylodyn_b200 import genealogy as gen


@pytest.mark.parametrize("name", ["changepoint", "constant"])
def test_coal_lik_init_reproduces_cached_init(name, request):
    g = request.getfixturevalue(name)
    a = {k: g.setting("Init.args." + k) for k in ("samp_times", "n_sampled", "coal_times")}
    grid = g.times[::g.gridsize]
    ini = gen.coal_lik_init(a["samp_times"], a["n_sampled"], a["coal_times"], grid)
    assert ini["ng"] == int(g.init["ng"][0])
    for k in ("t", "l", "C", "D", "y", "gridrep"):
        assert np.array_equal(ini[k], g.init[k]), k
    assert np.array_equal(ini["args"]["grid"], g.setting("Init.args.grid"))
    assert ini["gridrep"][-1] == 0 and ini["gridrep"].sum() == len(ini["C"])


def test_summarize_phylo_ebola():
    z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
    s = gen.summarize_phylo(z["edge"], z["edge_length"], int(z["ntip"]))
    # anchors from SURVEY.md Appendix D
    # phylodyn is not in the reference tree (parity unpinned): exact unique() of the tip heights gives 726
    # sampling times; SURVEY.md quotes 599 after rounding the 1e-6-level jitter of the stored branch lengths.
    assert s["n_sampled"].sum() == 1010 and 599 <= len(s["samp_times"]) <= 1010
    assert s["samp_times"][0] == 0.0 and s["samp_times"][-1] == pytest.approx(1.30137, abs=1e-5)
    assert len(s["coal_times"]) == 1009
    assert s["coal_times"][0] == pytest.approx(0.016881, abs=1e-6)
    assert s["coal_times"][-1] == pytest.approx(1.358504, abs=1e-6)
    times = np.linspace(0, 1.36, 2001)
    ini = gen.coal_lik_init(s["samp_times"], s["n_sampled"], s["coal_times"], times[::50])
    assert ini["gridrep"].sum() == len(ini["C"]) and ini["ng"] == 41
    assert ini["y"].sum() == 1009


def test_volz_thin_sir_synthetic_genealogy():
    """configs[4]-style synthetic genealogy: seeded, valid for coal_lik_init."""
    from oracle import oracle as orc
    times = np.linspace(0, 100, 2001)
    ode = orc.ODE_rk45([1e6, 1.0], times, [2.2, 0.2], [1e6], [0, 2, 0, 1])
    betaN = orc.betaTs([2.2, 0.2], times, [1e6], [0, 2, 0, 1])
    from oracle import sampler
    rng = np.random.default_rng(20260)
    c = sampler.volz_thin_sir(ode, 90.0, [0, 10, 20, 40, 80, 85], [100, 100, 145, 145, 9, 1], betaN, rng)
    assert len(c["coal_times"]) == 499
    assert np.all(np.diff(c["coal_times"]) >= 0) and c["coal_times"].max() <= 90.0
    ini = gen.coal_lik_init(c["samp_times"], c["n_sampled"], c["coal_times"], times[::50])
    assert ini["gridrep"].sum() == len(ini["C"]) and ini["y"].sum() == 499


def test_rdata_reader_on_the_reference_data_file():
    """lnaphylodyn_b200.rdata reads the reference's own `data/Ebola_Sierra_Leone2014.RData` (where the reference tree is mounted)
    into the arrays the committed fixture holds; without the reference tree, the reader is exercised on an RDX2 stream built
    here (a list with names, integer matrix with dim, real and character vectors)."""
    import gzip
    import struct
    import tempfile
    from lnaphylodyn_b200 import rdata
    path = "/root/reference/data/Ebola_Sierra_Leone2014.RData"
    z = np.load(os.path.join(GOLDEN, "ebola_tree.npz"))
    if os.path.exists(path):
        t = rdata.read_phylo(path)
        assert np.array_equal(t["edge"], z["edge"]) and np.array_equal(t["edge_length"], z["edge_length"])
        assert t["ntip"] == int(z["ntip"]) and t["Nnode"] == int(z["Nnode"][0])
        s = gen.summarize_phylo(t["edge"], t["edge_length"], t["ntip"])
        assert len(s["coal_times"]) == t["ntip"] - 1
    # a hand-built RDX2 file: pairlist(tree = list(edge = int[2x2], edge.length = dbl[2], tip.label = chr[2], Nnode = 1L))
    be_i, be_d = lambda *v: struct.pack(">%di" % len(v), *v), lambda *v: struct.pack(">%dd" % len(v), *v)
    def chars(s):
        return be_i(9 | (1 << 15), len(s)) + s.encode()
    def sym(s):
        return be_i(1) + chars(s)
    def strvec(v):
        return be_i(16, len(v)) + b"".join(chars(x) for x in v)
    names = be_i(2 | (1 << 10)) + sym("names") + strvec(["edge", "edge.length", "tip.label", "Nnode"]) + be_i(254)
    edge = be_i(13 | (1 << 9), 4, 3, 3, 1, 2) + be_i(2 | (1 << 10)) + sym("dim") + be_i(13, 2, 2, 2) + be_i(254)
    body = be_i(19 | (1 << 9), 4) + edge + be_i(14, 2) + be_d(0.5, 0.25) + strvec(["a", "b"]) + be_i(13, 1, 1) + names
    stream = b"RDX2\nX\n" + be_i(2, 0x030602, 0x020300) + be_i(2 | (1 << 10)) + sym("tree") + body + be_i(254)
    with tempfile.NamedTemporaryFile(suffix=".RData", delete=False) as f:
        f.write(gzip.compress(stream))
    try:
        t = rdata.read_phylo(f.name)
    finally:
        os.unlink(f.name)
    assert t["edge"].tolist() == [[3, 1], [3, 2]] and t["edge_length"].tolist() == [0.5, 0.25] and t["ntip"] == 2 and t["Nnode"] == 1
