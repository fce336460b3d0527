"""N>1 path on the CPU: two gloo ranks deal requests round-robin, "generate" deterministic frames
of ragged lengths, and rank 0 gathers them in request order (no collective on the decode path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sesameai.parallel import gather_frames, run_sharded, shard_requests


def _fake_generate(r: int) -> torch.Tensor:
    n = 3 + (r * 5) % 7  # ragged: utterances end at different EOS frames
    return (torch.arange(n * 32, dtype=torch.int32).view(n, 32) + 1000 * r) % 2051


def _worker(rank, world, port, n_requests, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = run_sharded(n_requests, _fake_generate, rank, world)
        assert sorted(local) == shard_requests(n_requests, rank, world)
        out = gather_frames(local, n_requests, rank, world, max_frames=16)
        if rank == 0:
            ok = len(out) == n_requests and all(torch.equal(out[r], _fake_generate(r)) for r in range(n_requests))
            q.put(ok)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_shard_requests_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(r for k in range(world) for r in shard_requests(37, k, world))
        assert seen == list(range(37))


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_single_rank_gather_needs_no_process_group():
    local = run_sharded(3, _fake_generate, 0, 1)
    out = gather_frames(local, 3, 0, 1)
    assert all(torch.equal(out[r], _fake_generate(r)) for r in range(3))
