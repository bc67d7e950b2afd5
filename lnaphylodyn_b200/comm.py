"""Multi-GPU plumbing behind the C ABI: the NCCL all-gather of the per-chain summary rows
(`lna_comm_*`, `lna_chains_allgather_summaries`) and the single-process multi-device chain driver (`lna_multi_*`).

One process per GPU (torchrun): `Comm.from_torch_distributed()` distributes NCCL's unique id with one
torch.distributed broadcast and creates the communicator inside the library; after that no Python-level collective
is involved in the exchange.  One process for all GPUs (what an R session is): `MultiChains`.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import LnaError, check
from .api import _d, _ptr


def _parse_cpulist(txt):
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_device_numa(pci_domain, pci_bus, pci_device, sysfs="/sys"):
    """One process per GPU: keep this process (and so the pinned host buffers it is about to allocate and first-touch) on the
    NUMA node the GPU hangs off.  Eight ranks streaming proposals to eight GPUs otherwise share whichever memory controllers
    and PCIe root complexes the scheduler happened to put them on.  Returns a dict describing what was done; a host without
    NUMA information (numa_node = -1, no sysfs) is left alone."""
    import os
    info = {"numa_node": None, "bound": False}
    try:
        dev = f"{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0"
        node = int(open(os.path.join(sysfs, "bus/pci/devices", dev, "numa_node")).read())
        info.update(pci=dev, numa_node=node)
        if node < 0:
            return info
        cpus = _parse_cpulist(open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")).read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return info
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed))
    except (OSError, ValueError, AttributeError) as e:
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info


class Comm:
    """Wraps lna_comm (an NCCL communicator owned by the library)."""

    def __init__(self, nranks, rank, unique_id, device):
        L = _lib.lib()
        self._id = bytes(unique_id)
        buf = C.create_string_buffer(self._id, 128)
        self.h = L.lna_comm_create(int(nranks), int(rank), buf, int(device))
        if not self.h:
            raise LnaError("lna_comm_create: " + L.lna_last_error().decode())
        self.nranks, self.rank, self.device = int(nranks), int(rank), int(device)
        self._fin = weakref.finalize(self, L.lna_comm_destroy, self.h)

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(_lib.lib().lna_comm_unique_id(buf), "lna_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device, group=None):
        """Rank 0 draws the id, one broadcast hands it to the other ranks (any launcher-side channel would do)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return cls(world, rank, bytes(t.cpu().numpy().tobytes()), device)

    def allgather_summaries(self, chains, B_total, out_ptr):
        """All B_total x 4 summary rows into device memory at `out_ptr` (asynchronous on the problem's stream)."""
        check(_lib.lib().lna_chains_allgather_summaries(chains.h, self.h, int(B_total), out_ptr), "lna_chains_allgather_summaries")


def bench_allgather(cx, pb, n_chains, reps=50):
    """bench.py leg: time the summary all-gather of `n_chains` chains per rank behind the C ABI (CUDA events on the problem's
    stream, max over ranks) and check it against the chains' own summaries and torch's all_gather_into_tensor."""
    import torch
    import torch.distributed as dist
    from .chains import Chains
    from conftest import Golden
    g = Golden("changepoint_sim")
    par0 = np.concatenate([[g.x_r[0], 1.0], g.setting("control.R0"), g.setting("control.gamma"), g.setting("control.ch"),
                           g.setting("control.hyper")])
    ch = Chains(pb, np.tile(par0, (n_chains, 1)), g.setting("control.traj"))
    from .chains import vignette_opts
    for it in range(3):
        ch.step_rng(vignette_opts(g, update_traj=it > 0), 3, it, cx.rank * n_chains)
    ch.sync()
    comm = Comm.from_torch_distributed(cx.local_rank)
    total = n_chains * cx.world
    out = torch.zeros((total, 4), dtype=torch.float64, device=cx.dev)
    for _ in range(5):
        comm.allgather_summaries(ch, total, out.data_ptr())
    torch.cuda.synchronize(cx.dev)
    cx.barrier()
    e0, e1 = cx.events()
    e0.record(cx.stream)
    for _ in range(reps):
        comm.allgather_summaries(ch, total, out.data_ptr())
    e1.record(cx.stream)
    torch.cuda.synchronize(cx.dev)
    us, = cx.max_over_ranks(e0.elapsed_time(e1) * 1e3 / reps)

    class _Dev:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (int(ptr), False), "version": 2}

    local = torch.as_tensor(_Dev(ch.summary_ptr(), (n_chains, 4)), device=cx.dev)
    ref = torch.empty((total, 4), dtype=torch.float64, device=cx.dev)
    dist.all_gather_into_tensor(ref, local.contiguous())
    torch.cuda.synchronize(cx.dev)
    ok = bool(torch.equal(ref, out)) and bool(torch.isfinite(out[:, 0]).all())
    return {"us_per_allgather": us, "rows": total, "bytes": total * 32, "ranks": cx.world, "matches_torch_all_gather": ok,
            "call": "lna_chains_allgather_summaries: one ncclAllGather on preallocated buffers behind the C ABI, no size "
                    "exchange, no host synchronisation", "nccl_version": _lib.lib().lna_comm_available()}


class MultiChains:
    """B chains sharded over several GPUs and driven by THIS process (wraps lna_multi).  Same observation interface as
    chains.Chains: step_rng / step_mcmc2_rng / get / history."""

    def __init__(self, cfg_problem_args, par, OriginTraj, devices, free_running=True):
        """cfg_problem_args: dict(times, x_r, x_i, gridsize, t_correct, init) as for api.Problem."""
        from . import api
        a = cfg_problem_args
        # a throw-away Problem on the first device builds the cfg/init structs and validates them
        self._pb0 = api.Problem(a["times"], a["x_r"], a["x_i"], a["gridsize"], a["t_correct"], a["init"], device=int(devices[0]))
        pb = self._pb0
        par = _d(par)
        self.B, self.npar = par.shape[0], 2 + pb.npt
        self.p, self.k, self.nrow, self.npt = pb.p, pb.k, pb.nrow, pb.npt
        self.n_norm = (self.npar - 1) + pb.k * pb.p
        cfg = _lib.LnaCfg(0, 0, pb.x_i[0], pb.x_i[1], pb.x_i[2], pb.x_i[3], pb.gridsize, len(pb.times),
                          pb.times.ctypes.data_as(_lib._dp), pb.x_r.ctypes.data_as(_lib._dp), pb.t_correct)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        L = _lib.lib()
        self.h = L.lna_multi_create(C.byref(cfg), C.byref(pb._ini), len(devices), devs, self.B)
        if not self.h:
            raise LnaError("lna_multi_create: " + L.lna_last_error().decode())
        self._fin = weakref.finalize(self, L.lna_multi_destroy, self.h)
        self.devices = [int(d) for d in devices]
        E, _ = pb._mats(OriginTraj, (pb.k, pb.p))
        if E.shape[0] == 1 and self.B > 1:
            E = np.ascontiguousarray(np.broadcast_to(E, (self.B, E.shape[1])))
        check(L.lna_multi_init(self.h, _ptr(par), _ptr(E)), "lna_multi_init")
        if free_running:
            check(L.lna_multi_set_free_running(self.h, 1), "lna_multi_set_free_running")
        self.pb = pb  # (for code that reads dimensions off chains.pb)

    def step_rng(self, opts, seed, iteration, chain_offset=0):
        check(_lib.lib().lna_multi_step_rng(self.h, C.byref(opts), int(seed), int(iteration), int(chain_offset)), "lna_multi_step_rng")

    def step_mcmc2_rng(self, opts, seed, iteration, chain_offset=0):
        check(_lib.lib().lna_multi_step_mcmc2_rng(self.h, C.byref(opts), int(seed), int(iteration), int(chain_offset)),
              "lna_multi_step_mcmc2_rng")

    def set_joint_transform(self, T):
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.float64), (self.B,) + np.shape(T)[-2:]))
        check(_lib.lib().lna_multi_set_joint_transform(self.h, int(T.shape[-1]), _ptr(T)), "lna_multi_set_joint_transform")

    def sync(self):
        check(_lib.lib().lna_multi_sync(self.h), "lna_multi_sync")

    def get(self):
        pb, B = self._pb0, self.B
        par, org = np.empty((B, self.npar)), np.empty((B, pb.k * pb.p))
        lat = np.empty((B, pb.nrow * (pb.p + 1)))
        cl, lo, cnt = np.empty(B), np.empty(B), np.zeros((B, 4), dtype=np.int32)
        check(_lib.lib().lna_multi_get(self.h, _ptr(par), _ptr(org), _ptr(lat), _ptr(cl), _ptr(lo), _ptr(cnt)), "lna_multi_get")
        return {"par": par, "OriginTraj": pb._unmats(org, (pb.k, pb.p), False),
                "LatentTraj": pb._unmats(lat, (pb.nrow, pb.p + 1), False), "coalLog": cl, "logOrigin": lo,
                "n_full_evals": cnt[:, 0], "n_cheap_evals": cnt[:, 1], "n_mh_accept": cnt[:, 2], "status": cnt[:, 3]}

    def latent(self):
        pb, B = self._pb0, self.B
        lat = np.empty((B, pb.nrow * (pb.p + 1)))
        check(_lib.lib().lna_multi_get(self.h, None, None, _ptr(lat), None, None, None), "lna_multi_get")
        return pb._unmats(lat, (pb.nrow, pb.p + 1), False)

    def history_begin(self, niter):
        self._hist_niter = int(niter)
        check(_lib.lib().lna_multi_history_begin(self.h, int(niter)), "lna_multi_history_begin")

    def history_record(self, i):
        check(_lib.lib().lna_multi_history_record(self.h, int(i)), "lna_multi_history_record")

    def history_get(self, par, l, l1):
        check(_lib.lib().lna_multi_history_get(self.h, self._hist_niter, _ptr(par), _ptr(l), _ptr(l1)), "lna_multi_history_get")

    def gather_summaries(self, use_nccl=None):
        """(B, 4) summary rows of all chains; with NCCL every device also ends up holding them (lna_multi_gathered_dev)."""
        if use_nccl is None:
            use_nccl = len(set(self.devices)) == len(self.devices) and len(self.devices) > 1
        out = np.empty((self.B, 4))
        check(_lib.lib().lna_multi_gather_summaries(self.h, int(bool(use_nccl)), _ptr(out)), "lna_multi_gather_summaries")
        return out
