"""The N>1 host path on CPU: env-axis sharding, the one gather of the controls and the max-over-ranks timing
reduction, run as two gloo processes (the GPU path uses the same helpers over NCCL; there is no collective on the
solve path itself: SURVEY.md 8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpx_b200.utils.sharding import env_slice, gather_controls, max_over_ranks


def test_env_slice_partitions_the_env_axis():
    for total in (0, 1, 7, 1024, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [env_slice(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        env_slice(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, nj, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = env_slice(total, world, rank)
        # stand-in for the per-rank solve: the control of env e is a deterministic function of e only
        env = torch.arange(lo, hi, dtype=torch.float32)
        tau = env[:, None] * 10.0 + torch.arange(nj, dtype=torch.float32)[None, :]
        full = gather_controls(tau, n_env_total=total, dst=0)
        ms = max_over_ranks(1.0 + rank)
        if rank == 0:
            want = torch.arange(total, dtype=torch.float32)[:, None] * 10.0 + torch.arange(nj, dtype=torch.float32)[None, :]
            q.put((bool(torch.equal(full, want)), tuple(full.shape), ms))
        else:
            q.put((full is None, None, ms))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_controls_two_ranks_gloo(total):
    world, nj = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, nj, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res)
    assert any(r[1] == (total, nj) for r in res)
    assert all(r[2] == 2.0 for r in res)  # max over ranks of (1 + rank)


def test_single_process_is_identity():
    tau = torch.randn(5, 3)
    assert gather_controls(tau) is tau
    assert max_over_ranks(3.5) == 3.5
