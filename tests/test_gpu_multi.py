"""Multi-GPU behind the boundary: the single-process multi-device driver (lna_multi) and the NCCL
all-gather of the summary rows (lna_comm).  On a one-GPU box the sharding logic runs with the same device listed twice
(two problems, two resident batches, two sets of streams) and NCCL with a one-rank communicator; with >= 2 GPUs the same
tests run over distinct devices and a real ncclAllGather."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _par0(g):
    return np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])


def _args(g):
    return {"times": g.times, "x_r": g.x_r, "x_i": g.x_i, "gridsize": g.gridsize, "t_correct": g.t_correct, "init": g.init}


@pytest.mark.parametrize("shards", [2, 3])
def test_sharded_chains_equal_single_batch(changepoint, shards):
    """Chains are independent and the Philox streams are keyed by the GLOBAL chain index: a batch sharded over `shards`
    device slots (distinct GPUs when the box has them, else the same GPU several times) reproduces the single batch per
    chain, bit for bit -- state, counters, per-iteration history and the gathered summary rows.  Odd batch size: ragged shards."""
    from lnaphylodyn_b200 import _lib, api, chains as chm, comm
    g = changepoint
    ndev = _lib.lib().lna_device_count()
    devices = [i % ndev for i in range(shards)]
    B, iters, seed = 37, 5, 21
    rng = np.random.default_rng(0)
    par = np.tile(_par0(g), (B, 1))
    par[:, 2] *= np.exp(0.02 * rng.standard_normal(B))
    eps = 0.1 * rng.standard_normal((B, 40, 2))
    single = chm.Chains(api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init), par, eps)
    multi = comm.MultiChains(_args(g), par, eps, devices)
    single.history_begin(iters)
    multi.history_begin(iters)
    for it in range(iters):
        opts = chm.vignette_opts(g, update_traj=it >= 2)
        single.step_rng(opts, seed, it)
        multi.step_rng(opts, seed, it)
        single.history_record(it)
        multi.history_record(it)
    a, b = single.get(), multi.get()
    for k in ("par", "OriginTraj", "LatentTraj", "coalLog", "logOrigin", "n_full_evals", "n_cheap_evals", "n_mh_accept", "status"):
        assert np.array_equal(a[k], b[k]), k
    ha = [np.empty((B, iters, 44)), np.empty((B, iters)), np.empty((B, iters))]
    hb = [np.empty((B, iters, 44)), np.empty((B, iters)), np.empty((B, iters))]
    single.history_get(*ha)
    multi.history_get(*hb)
    for x, y in zip(ha, hb):
        assert np.array_equal(x, y)
    assert not np.array_equal(ha[0][:, 0], ha[0][:, -1])
    rows = multi.gather_summaries(use_nccl=len(set(devices)) == shards)
    assert np.array_equal(rows[:, 0], a["coalLog"]) and np.array_equal(rows[:, 1], a["logOrigin"])
    assert np.array_equal(rows[:, 2], a["n_full_evals"].astype(float))


def test_driver_over_device_slots(changepoint):
    """mcmc.General_MCMC_with_ESlice(devices=...) -- one process, several resident batches -- returns the run of the
    single-batch call."""
    from lnaphylodyn_b200 import _lib, mcmc
    from test_gpu_mcmc_driver import vignette_args
    g = changepoint
    coal, prior, proposal, control = vignette_args(g)
    nch = len(g.x_r) - 1
    ndev = _lib.lib().lna_device_count()
    kw = dict(gridsize=50, niter=6, burn=0, thin=2, changetime=g.x_r[1:], DEMS=("S", "I"), prior=prior, proposal=proposal,
              control=control, ESS_vec=(0, 1, 0, 0, 1, 1, 0), likelihood="volz", model="SIR", Index=(0, 1), nparam=2,
              options={"burn1": 3, "parIdlist": {"c": 4, "d": nch + 5}, "isLoglist": {"c": 1, "d": 0},
                       "priorIdlist": {"c": 3, "d": 5}, "hyper": False}, n_chains=9, seed=4)
    one = mcmc.General_MCMC_with_ESlice(coal, g.times, g.t_correct, g.x_r[0], **kw)
    two = mcmc.General_MCMC_with_ESlice(coal, g.times, g.t_correct, g.x_r[0], devices=[0, 1 % ndev], **kw)
    for k in ("par", "Trajectory", "l", "l1"):
        assert np.array_equal(one[k], two[k]), k
    assert np.array_equal(one["MCMC_obj"]["coalLog"], two["MCMC_obj"]["coalLog"])


def test_nccl_allgather_of_summaries_one_rank(changepoint):
    """lna_comm on this process's GPU (a one-rank NCCL communicator): the gathered block is the chains' own summary block;
    the multi-rank form runs in bench.py (`summary_allgather`, checked there against torch's all_gather_into_tensor)."""
    import torch
    from lnaphylodyn_b200 import _lib, api, chains as chm, comm
    g = changepoint
    assert _lib.lib().lna_comm_available() > 0, _lib.lib().lna_last_error()
    B = 300
    pb = api.Problem(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)
    ch = chm.Chains(pb, np.tile(_par0(g), (B, 1)), g.setting("control.traj"))
    for it in range(2):
        ch.step_rng(chm.vignette_opts(g, update_traj=True), 5, it)
    c = comm.Comm(1, 0, comm.Comm.unique_id(), 0)
    out = torch.zeros((B, 4), dtype=torch.float64, device="cuda:0")
    c.allgather_summaries(ch, B, out.data_ptr())
    pb.sync()
    s = ch.get()
    got = out.cpu().numpy()
    assert np.array_equal(got[:, 0], s["coalLog"]) and np.array_equal(got[:, 1], s["logOrigin"])
    assert np.array_equal(got[:, 2], s["n_full_evals"].astype(float))
    with pytest.raises(_lib.LnaError, match="shard"):
        c.allgather_summaries(ch, B + 1, out.data_ptr())
