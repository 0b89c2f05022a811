"""world_size-2 gloo test of the multi-GPU host logic (shard ranges + state all_gather)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jaxsim_b200.api as js
from jaxsim_b200 import distributed as D


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 8, 65536, 65537):
        for W in (1, 2, 3, 8):
            r = [D.shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)


def _fake_data(lo, hi, n, nc):
    B = hi - lo
    idx = torch.arange(lo, hi, dtype=torch.float64)[:, None]
    mk = lambda w, k: idx * 100 + k + torch.arange(w, dtype=torch.float64)[None, :] * 0.01  # noqa: E731
    return js.data.JaxSimModelData(
        _joint_positions=mk(n, 1), _joint_velocities=mk(n, 2), _base_quaternion=mk(4, 3),
        _base_linear_velocity=mk(3, 4), _base_angular_velocity=mk(3, 5), _base_position=mk(3, 6),
        contact_state={"tangential_deformation": mk(3 * nc, 7).reshape(B, nc, 3)},
    )


def _worker(rank, world, port, B, n, nc, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(B, rank, world)
        full = D.all_gather_state(_fake_data(lo, hi, n, nc))
        ref = D.pack_state(_fake_data(0, B, n, nc))
        ok = full.shape == ref.shape and torch.equal(full, ref)
        leaves = D.unpack_state(full, n, nc)
        ok = ok and leaves["tangential_deformation"].shape == (B, nc, 3) and leaves["_base_quaternion"].shape == (B, 4)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 9])
def test_all_gather_state_gloo_world2(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 5, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
