"""Utterance-level data parallelism: one process per GPU, each with its own model
replica and KV caches; requests are dealt round-robin, nothing is exchanged on the decode path, and
only the finished code frames (int32 [frames, 32], <= 16 KB per 10 s utterance) are gathered on
rank 0 at the end.  Works over any ``torch.distributed`` backend (NCCL on the GPU box, gloo in the
CPU tests); the reference itself has no multi-device path (batch 1, one process)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_requests(n_requests: int, rank: int, world: int) -> List[int]:
    """Indices of the requests rank ``rank`` serves (round-robin deal: request r -> GPU r mod world)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return list(range(rank, n_requests, world))


def run_sharded(n_requests: int, generate_fn: Callable[[int], torch.Tensor], rank: int, world: int) -> Dict[int, torch.Tensor]:
    """Run ``generate_fn(request_index) -> int32 [frames, 32]`` for this rank's shard."""
    return {r: generate_fn(r).to(torch.int32) for r in shard_requests(n_requests, rank, world)}


def gather_frames(local: Dict[int, torch.Tensor], n_requests: int, rank: int, world: int, device="cpu",
                  max_frames: int = 2048, codebooks: int = 32) -> Optional[List[torch.Tensor]]:
    """Final host gather: rank 0 returns the per-request frame tensors in request order, other ranks
    return None.  Utterances stop at different EOS frames, so lengths travel with the padded payload."""
    if world == 1:
        return [local[r].cpu() for r in range(n_requests)]
    per_rank = (n_requests + world - 1) // world
    pay = torch.zeros(per_rank, max_frames * codebooks + 2, dtype=torch.int32, device=device)
    pay[:, 0] = -1
    for slot, r in enumerate(shard_requests(n_requests, rank, world)):
        f = local[r]
        if f.shape[0] > max_frames:
            raise ValueError("utterance longer than max_frames")
        pay[slot, 0] = r
        pay[slot, 1] = f.shape[0]
        pay[slot, 2 : 2 + f.numel()] = f.reshape(-1).to(device)
    bufs = [torch.empty_like(pay) for _ in range(world)] if rank == 0 else None
    dist.gather(pay, bufs, dst=0)
    if rank != 0:
        return None
    out: List[Optional[torch.Tensor]] = [None] * n_requests
    for b in bufs:
        b = b.cpu()
        for row in b:
            r, n = int(row[0]), int(row[1])
            if r >= 0:
                out[r] = row[2 : 2 + n * codebooks].view(n, codebooks).clone()
    assert all(o is not None for o in out)
    return out  # type: ignore[return-value]


def serve_sharded(model, requests, rank: int, world: int, *, lanes: int, temperature: float, topk: int, groups: int = 1,
                  device="cpu", max_frames: int = 2048, codebooks: int = 32, server=None) -> Optional[List[torch.Tensor]]:
    """BASELINE config 5 ("N concurrent requests on 1/2/4/8 GPUs"): this rank serves its round-robin shard of
    ``requests`` (``sesameai.serving.Request``) with continuous batching -- ``groups`` lane groups of ``lanes``
    cache lanes each -- and the finished frames are gathered on rank 0.  No collective on the decode path.
    ``server``: a ``LaneGroups`` built beforehand (a serving process keeps its decode contexts alive)."""
    from .serving import LaneGroups

    mine = [requests[i] for i in shard_requests(len(requests), rank, world)]
    if server is None:
        server = LaneGroups(model, groups, lanes, temperature, topk, max_frames=max_frames)
    served = server.run(mine) if mine else {}
    local = {r.rid: served[r.rid].to(torch.int32) for r in mine}
    return gather_frames(local, len(requests), rank, world, device=device, max_frames=max_frames, codebooks=codebooks)
