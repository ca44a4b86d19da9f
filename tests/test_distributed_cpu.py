"""CPU, world_size 2, gloo: the multi-GPU plumbing of the path (batch sharding + the one all-gather of per-sample joints)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keypointfusion_b200.runtime import all_gather_joints, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_batch(n_items, rank, world)
    # each rank "computes" joints for its slice: value encodes the global sample id, so the gather order is checkable
    joints = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(hi - lo, 21, 3).contiguous()
    allj = all_gather_joints(joints)
    if rank == 0:
        out.put(allj[:, 0, 0].tolist())
    dist.destroy_process_group()


def test_shard_batch_partitions():
    for n, w in ((64, 2), (512, 8), (7, 2), (5, 8)):
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_all_gather_joints_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert got == [float(i) for i in range(8)]          # rank-major == global sample order


def test_all_gather_single_process_is_identity():
    j = torch.randn(4, 21, 3)
    assert all_gather_joints(j) is j
