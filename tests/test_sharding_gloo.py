"""world_size = 2 on CPU (gloo): the host-side logic of the multi-GPU path -- contiguous chain shards,
rank-independent replayable streams, ordered all-gather of the per-chain summaries."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lnaphylodyn_b200 import chains as chm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = chm.shard_range(B, rank, world)
    nrm, unf = chm.draw_streams(e - b, 7, seed=5, iteration=3, chain_offset=b)
    # a fake per-chain summary that encodes the global chain id
    local = torch.tensor(np.stack([np.arange(b, e), nrm[:, 0], unf[:, 0], np.full(e - b, rank)], axis=1))
    allsum = chm.gather_summaries(local, B_total=B)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), allsum.numpy())
    np.save(os.path.join(out_dir, f"n{rank}.npy"), nrm)
    dist.destroy_process_group()


def test_two_rank_shards_and_gather(tmp_path):
    B, world = 101, 2  # odd: shards differ by one chain
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    ref_n, ref_u = chm.draw_streams(B, 7, seed=5, iteration=3)
    got = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])                       # every rank sees the same gathered block
    g = got[0]
    assert g.shape == (B, 4)
    assert np.array_equal(g[:, 0], np.arange(B))                # global chain order
    assert np.array_equal(g[:, 1], ref_n[:, 0]) and np.array_equal(g[:, 2], ref_u[:, 0])  # rank-independent streams
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"n{r}.npy") for r in range(world)]), ref_n)
    assert (g[:, 3] == 0).sum() == 51 and (g[:, 3] == 1).sum() == 50


def test_bind_host_to_device_numa_on_a_fake_sysfs(tmp_path):
    """One process per GPU stays on its GPU's NUMA node (bench.py e2e legs at N > 1): the helper reads the PCI device's
    numa_node and the node's cpulist, intersects with the allowed CPUs, and leaves hosts without NUMA information alone."""
    import os
    from lnaphylodyn_b200 import comm
    pci = tmp_path / "bus/pci/devices/0000:1b:00.0"
    pci.mkdir(parents=True)
    (pci / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    before = os.sched_getaffinity(0)
    first = min(before)
    (node / "cpulist").write_text(f"{first},100000-100003\n")
    try:
        info = comm.bind_host_to_device_numa(0, 0x1b, 0, sysfs=str(tmp_path))
        assert info["bound"] and info["numa_node"] == 1 and os.sched_getaffinity(0) == {first}
    finally:
        os.sched_setaffinity(0, before)
    (pci / "numa_node").write_text("-1\n")
    assert not comm.bind_host_to_device_numa(0, 0x1b, 0, sysfs=str(tmp_path))["bound"]
    assert not comm.bind_host_to_device_numa(0, 0x1c, 0, sysfs=str(tmp_path))["bound"]   # unknown device: no-op, no raise
    assert os.sched_getaffinity(0) == before
    assert comm._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
