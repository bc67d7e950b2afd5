"""Resident chain batches: the batched equivalent of the reference's R-level MCMC loop
General_MCMC_with_ESlice (R/LNA_chpt_slice.R:723-849) over B independent chains on one GPU, and
the torch.distributed plumbing that shards chains across the GPUs of one box.

State lives in HBM between iterations (the R list MCMC_obj, general_change_point.R:558-561);
only per-chain summaries are all-gathered across ranks.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import LnaMcmcOpts, LnaError, check
from .api import ESLICE_MAX_UNIF, _d, _ptr

ITER_UNIF = 3 + 2 * ESLICE_MAX_UNIF
ITER_UNIF_GENERAL = 2 * 4 + 2 * ESLICE_MAX_UNIF  # with a general Update_Param_each_Norm entry list (opts.n_mh > 0)


def make_opts(prior, proposal, ESS_vec, update_traj=True, hyper_noop_mh=True, spec_width=0,
              gamma_prior_id=3, hyper_prior_id=5):
    """Translate the vignettes' `prior` / `proposal` lists (by POSITION, like the R code does:
    prlist = list(prior[[1]][1:2], prior[[2]], prior[[3]], prior[[5]]), LNA_chpt_slice.R:743)."""
    pr = [np.asarray(v, dtype=np.float64).ravel() for v in prior]
    po = [np.asarray(v, dtype=np.float64).ravel() for v in proposal]
    o = LnaMcmcOpts()
    pr8 = np.concatenate([pr[0][:2], pr[1][:2], pr[2][:2], pr[hyper_prior_id - 1][:2]])
    for i in range(8):
        o.prior8[i] = float(pr8[i])
    o.gamma_prior[0], o.gamma_prior[1] = float(pr[gamma_prior_id - 1][0]), float(pr[gamma_prior_id - 1][1])
    # Update_Param_each_Norm is called with proposeIdlist = priorIdlist (LNA_chpt_slice.R:791-792)
    o.gamma_prop = float(po[gamma_prior_id - 1][0])
    for i in range(7):
        o.ESS_vec[i] = int(ESS_vec[i])
    o.update_traj, o.hyper_noop_mh, o.spec_width = int(update_traj), int(hyper_noop_mh), int(spec_width)
    return o


def make_opts_general(prior, proposal, ESS_vec, entries, hyper=False, update_traj=True, spec_width=0, hyper_prior_id=5):
    """Like make_opts for an arbitrary Update_Param_each_Norm entry list (R/LNA_chpt_slice.R:169-265).  `entries`:
    (R's 1-based parId, isLog, priorId) per list entry, e.g. the Ebola vignette's ((1, 1, 1), (3, 1, 3), (nch + 5, 0, 5)).
    parId 1..4 = S0, I0, R0, gamma; the last par (nch + 5) = the hyper-parameter, which is the no-op entry with hyper = F and
    a log-normal random walk with hyper = T.  The proposal ids equal the prior ids (:791-792)."""
    pr = [np.asarray(v, dtype=np.float64).ravel() for v in prior]
    po = [np.asarray(v, dtype=np.float64).ravel() for v in proposal]
    o = make_opts(prior, proposal, ESS_vec, update_traj=update_traj, hyper_noop_mh=False, spec_width=spec_width,
                  hyper_prior_id=hyper_prior_id)
    if not 1 <= len(entries) <= 4:
        raise ValueError("1..4 Update_Param_each_Norm entries")
    o.n_mh = len(entries)
    for e, (par_id, is_log, pid) in enumerate(entries):
        if pid == 5:
            o.mh_kind[e], o.mh_index[e], o.mh_is_log[e] = (2 if hyper else 1), -1, 0
        elif 1 <= par_id <= 4 and 1 <= pid <= 4:
            o.mh_kind[e], o.mh_index[e], o.mh_is_log[e] = 0, int(par_id) - 1, int(is_log)
        else:
            raise ValueError(f"entry {(par_id, is_log, pid)}: parId must be 1..4 with priorId 1..4, or the hyper-parameter with priorId 5")
        o.mh_prior[e][0], o.mh_prior[e][1] = float(pr[pid - 1][0]), float(pr[pid - 1][1])
        o.mh_prop[e] = float(po[pid - 1][0])
    return o


class Chains:
    """B resident chains on one GPU (wraps lna_chains)."""

    def __init__(self, problem, par, OriginTraj, free_running=False):
        """free_running: the stream groups of a large batch are not joined after every step (lna_chains_set_free_running);
        get() / join() / sync() are the observation points."""
        self.pb = problem
        par = _d(par)
        self.B = par.shape[0]
        self.npar = 2 + problem.npt
        self.n_norm = (self.npar - 1) + problem.k * problem.p
        L = _lib.lib()
        self.h = L.lna_chains_create(problem.h, self.B)
        if not self.h:
            raise LnaError("lna_chains_create: " + L.lna_last_error().decode())
        self._fin = weakref.finalize(self, L.lna_chains_destroy, self.h)
        E, _ = problem._mats(OriginTraj, (problem.k, problem.p))
        if E.shape[0] == 1 and self.B > 1:
            E = np.ascontiguousarray(np.broadcast_to(E, (self.B, E.shape[1])))
        check(L.lna_chains_init(self.h, _ptr(par), _ptr(E)), "lna_chains_init")
        if free_running:
            check(L.lna_chains_set_free_running(self.h, 1), "lna_chains_set_free_running")

    def set_free_running(self, on=True):
        check(_lib.lib().lna_chains_set_free_running(self.h, int(bool(on))), "lna_chains_set_free_running")

    def join(self):
        """Order the problem's stream after everything enqueued on the chains' stream groups so far."""
        check(_lib.lib().lna_chains_join(self.h), "lna_chains_join")

    def sync(self):
        """Host waits for all stream groups."""
        check(_lib.lib().lna_chains_sync(self.h), "lna_chains_sync")

    @property
    def groups(self):
        return _lib.lib().lna_chains_groups(self.h)

    def step(self, opts, normals, uniforms):
        """One MCMC iteration. normals (B, npar-1 + k*p), uniforms (B, 47): host arrays."""
        n, u = _d(normals), _d(uniforms)
        nu = ITER_UNIF_GENERAL if opts.n_mh > 0 else ITER_UNIF
        assert n.shape == (self.B, self.n_norm) and u.shape == (self.B, nu), (n.shape, u.shape)
        check(_lib.lib().lna_chains_step(self.h, C.byref(opts), _ptr(n), _ptr(u)), "lna_chains_step")

    def step_dev(self, opts, normals_ptr, uniforms_ptr):
        """Same with streams already in HBM (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        check(_lib.lib().lna_chains_step_dev(self.h, C.byref(opts), normals_ptr, uniforms_ptr), "lna_chains_step_dev")

    def step_rng(self, opts, seed, iteration, chain_offset=0):
        """One iteration on streams drawn in HBM (Philox4x32-10; replayable with device_streams)."""
        check(_lib.lib().lna_chains_step_rng(self.h, C.byref(opts), int(seed), int(iteration), int(chain_offset)),
              "lna_chains_step_rng")

    def device_streams(self, seed, iteration, chain_offset=0, n_unif=ITER_UNIF):
        """The streams step_rng(seed, iteration, chain_offset) consumes, copied to the host (n_unif = ITER_UNIF_GENERAL for
        options with a general entry list)."""
        nrm, unf = np.empty((self.B, self.n_norm)), np.empty((self.B, n_unif))
        check(_lib.lib().lna_draw_streams(self.pb.h, self.B, int(chain_offset), self.n_norm, n_unif, int(seed),
                                          int(iteration), _ptr(nrm), _ptr(unf)), "lna_draw_streams")
        return nrm, unf

    def get(self):
        pb, B = self.pb, self.B
        par = np.empty((B, self.npar))
        org = np.empty((B, pb.k * pb.p))
        lat = np.empty((B, pb.nrow * (pb.p + 1)))
        cl, lo = np.empty(B), np.empty(B)
        cnt = np.zeros((B, 4), dtype=np.int32)
        check(_lib.lib().lna_chains_get(self.h, _ptr(par), _ptr(org), _ptr(lat), _ptr(cl), _ptr(lo), _ptr(cnt)),
              "lna_chains_get")
        return {"par": par, "OriginTraj": pb._unmats(org, (pb.k, pb.p), False),
                "LatentTraj": pb._unmats(lat, (pb.nrow, pb.p + 1), False), "coalLog": cl, "logOrigin": lo,
                "n_full_evals": cnt[:, 0], "n_cheap_evals": cnt[:, 1], "n_mh_accept": cnt[:, 2], "status": cnt[:, 3]}

    def step_mcmc2_rng(self, opts, seed, iteration, chain_offset=0):
        """One General_MCMC2 iteration on device-drawn streams."""
        check(_lib.lib().lna_chains_step_mcmc2_rng(self.h, C.byref(opts), int(seed), int(iteration), int(chain_offset)),
              "lna_chains_step_mcmc2_rng")

    def latent(self):
        pb, B = self.pb, self.B
        lat = np.empty((B, pb.nrow * (pb.p + 1)))
        check(_lib.lib().lna_chains_get(self.h, None, None, _ptr(lat), None, None, None), "lna_chains_get")
        return pb._unmats(lat, (pb.nrow, pb.p + 1), False)

    def history_begin(self, niter):
        check(_lib.lib().lna_chains_history_begin(self.h, int(niter)), "lna_chains_history_begin")

    def history_record(self, i):
        check(_lib.lib().lna_chains_history_record(self.h, int(i)), "lna_chains_history_record")

    def history_get(self, par, l, l1):
        check(_lib.lib().lna_chains_history_get(self.h, _ptr(par), _ptr(l), _ptr(l1)), "lna_chains_history_get")

    def exec_counts(self):
        """(FULL evaluations executed incl. speculative ones, warp-evaluations executed) since creation."""
        e, w = C.c_uint64(), C.c_uint64()
        check(_lib.lib().lna_chains_exec_counts(self.h, C.byref(e), C.byref(w)), "lna_chains_exec_counts")
        return int(e.value), int(w.value)

    def summary_ptr(self):
        self.join()
        return _lib.lib().lna_chains_summary_dev(self.h)


def draw_streams(B, n_norm, seed, iteration, chain_offset=0):
    """Replayable per-chain streams: chain c of a job draws from default_rng([seed, c, iteration])."""
    nrm = np.empty((B, n_norm))
    unf = np.empty((B, ITER_UNIF))
    for c in range(B):
        rng = np.random.default_rng([seed, chain_offset + c, iteration])
        nrm[c] = rng.standard_normal(n_norm)
        u = rng.random(ITER_UNIF)
        unf[c] = np.where(u == 0.0, 0.5, u)  # open interval like unif_rand()
    return nrm, unf


def vignette_opts(g, update_traj=True, spec_width=0):
    """Options of the Changpoint_sim vignette call (vignettes/Changpoint_sim.Rmd:89-98) from a golden cache."""
    prior = [g.setting("prior." + k) for k in ("pop_pr", "R0_pr", "gamma_pr", "mu_pr", "hyper_pr")]
    proposal = [g.setting("proposal." + k) for k in ("pop_prop", "R0_prop", "gamma_prop", "mu_prop", "hyper_prop")]
    ess = [0, 1, 0, 0, 1, 1, 0]
    if not update_traj:
        ess[4] = 0  # ESS_temp for i < burn1 (LNA_chpt_slice.R:773-774)
    return make_opts(prior, proposal, ess, update_traj=update_traj, hyper_noop_mh=True, spec_width=spec_width)


# ------------------------------------------------------------------------------------------------
# multi-GPU plumbing: contiguous chain shards, no data-path collective
# ------------------------------------------------------------------------------------------------

def shard_range(B, rank, world):
    """Contiguous block [begin, end) of the B global chains owned by `rank` (lna_shard)."""
    b, e = C.c_int64(), C.c_int64()
    check(_lib.lib().lna_shard(int(B), int(rank), int(world), C.byref(b), C.byref(e)), "lna_shard")
    return b.value, e.value


def gather_summaries(local, B_total=None, group=None):
    """All-gather the per-chain summary rows (coalLog, logOrigin, n_full_evals, status) of every rank, in global chain
    order, through torch.distributed (NCCL on device tensors, gloo on CPU tensors: the host-logic tests).  ONE
    all_gather_into_tensor on a buffer padded to the largest shard; the shard sizes come from lna_shard(B_total), not from
    an exchange.  B_total defaults to equal shards.  (Inside the library the same exchange is lna_chains_allgather_summaries
    over ncclAllGather: comm.py.)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n, cols = local.shape
    if B_total is None:
        B_total = n * world
    sizes = [shard_range(B_total, r, world) for r in range(world)]
    sizes = [e - b for b, e in sizes]
    if sizes[dist.get_rank(group)] != n:
        raise LnaError(f"gather_summaries: this rank holds {n} rows, its shard of {B_total} is {sizes[dist.get_rank(group)]}")
    mx = max(sizes)
    if mx == n and min(sizes) == mx:
        out = torch.empty((world * n, cols), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((mx, cols), dtype=local.dtype, device=local.device)
    pad[:n] = local
    out = torch.empty((world * mx, cols), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx:r * mx + s] for r, s in enumerate(sizes)], dim=0)


# ------------------------------------------------------------------------------------------------
# General_MCMC2 (constant_sim vignette) on resident chains
# ------------------------------------------------------------------------------------------------
ITER2_UNIF = 2 * 4 + 2 + 2 * ESLICE_MAX_UNIF


def make_mcmc2_opts(prior, proposal, entries=((1, 1, 1), (2, 0, 2), (3, 1, 3)), update_hyper=True, update_chpt=True,
                    update_traj=True, spec_width=0):
    """`entries`: (0-based par index, isLog, priorId) per Update_Param_each entry -- the vignette passes
    parIdlist = (2, 3, 4), isLoglist = (1, 0, 1), priorIdlist = (1, 2, 3) (vignettes/constant_sim.Rmd:87-96),
    and the proposal ids equal the prior ids (R/LNA_chpt_slice.R:611-612).  priorId 2 is dunif, the others dnorm
    (R/LNA_chpt_slice.R:306-320); update_hyper uses dgamma(hyper_pr) and proposal$hyper_prop."""
    from ._lib import LnaMcmc2Opts
    pr = [np.asarray(v, dtype=np.float64).ravel() for v in prior]
    po = [np.asarray(v, dtype=np.float64).ravel() for v in proposal]
    o = LnaMcmc2Opts()
    o.n_mh = len(entries)
    for e, (idx, is_log, pid) in enumerate(entries):
        o.mh_index[e], o.mh_is_log[e], o.mh_prior_kind[e] = int(idx), int(is_log), 1 if pid == 2 else 0
        o.mh_prior[e][0], o.mh_prior[e][1] = float(pr[pid - 1][0]), float(pr[pid - 1][1])
        o.mh_prop[e] = float(po[pid - 1][0])
    o.update_hyper, o.update_chpt, o.update_traj = int(update_hyper), int(update_chpt), int(update_traj)
    o.hyper_prop = float(po[4][0])
    o.hyper_prior[0], o.hyper_prior[1] = float(pr[4][0]), float(pr[4][1])
    o.spec_width = int(spec_width)
    return o


def make_joint_opts(prior, entries, update_chpt=True, update_traj=True, spec_width=0):
    """General_MCMC2's stage i >= warmup2: update_Param_joint(method = "admcmc") (R/LNA_chpt_slice.R:9-88) + change-point and
    trajectory slice updates.  `entries`: (joint index, isLog, priorId) per jointly proposed parameter, joint index 1 = I0,
    2 = R0, 3 = gamma, 4 = hyper (last); priorId 2 is dunif, 5 dgamma, the others dnorm (:44-58)."""
    from ._lib import LnaMcmc2Opts
    pr = [np.asarray(v, dtype=np.float64).ravel() for v in prior]
    o = LnaMcmc2Opts()
    o.joint, o.joint_d = 1, len(entries)
    for e, (idx, is_log, pid) in enumerate(entries):
        o.joint_index[e], o.joint_is_log[e] = int(idx), int(is_log)
        o.joint_prior_kind[e] = 1 if pid == 2 else 2 if pid == 5 else 0
        o.joint_prior[e][0], o.joint_prior[e][1] = float(pr[pid - 1][0]), float(pr[pid - 1][1])
    o.update_chpt, o.update_traj, o.spec_width = int(update_chpt), int(update_traj), int(spec_width)
    return o


def mvrnorm_transform(pcov):
    """T with raw' = raw + T z, z ~ N(0, I): MASS::mvrnorm's eS$vectors %*% diag(sqrt(pmax(ev, 0))) (R/LNA_chpt_slice.R:39).
    pcov: (..., d, d).  (Eigenvector signs / order are LAPACK's: the proposal DISTRIBUTION is the reference's, its
    realisation for given normals is not pinned -- MASS is a third-party package outside the reference tree.)"""
    ev, vec = np.linalg.eigh(np.asarray(pcov, dtype=np.float64))
    return vec * np.sqrt(np.maximum(ev, 0.0))[..., None, :]


def _set_joint_transform(self, T):
    T = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.float64), (self.B,) + np.shape(T)[-2:]))
    check(_lib.lib().lna_chains_set_joint_transform(self.h, int(T.shape[-1]), _ptr(T)), "lna_chains_set_joint_transform")


Chains.set_joint_transform = _set_joint_transform


def _step_mcmc2(self, opts, normals, uniforms):
    """One General_MCMC2 iteration. normals (B, nch + k*p [+ 4 with opts.joint]), uniforms (B, 54): host arrays."""
    n, u = _d(normals), _d(uniforms)
    assert u.shape == (self.B, ITER2_UNIF) and n.shape[0] == self.B, (n.shape, u.shape)
    check(_lib.lib().lna_chains_step_mcmc2(self.h, C.byref(opts), _ptr(n), _ptr(u)), "lna_chains_step_mcmc2")


Chains.step_mcmc2 = _step_mcmc2
