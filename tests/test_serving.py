"""Continuous batching (sesameai.serving) on the CPU with a scripted stand-in for the model: the scheduler must
give every request exactly the frames the reference loop (generator.py:283-294) would have produced for it
alone -- same feedback of the sampled frame, same per-stream EOS rule, same frame budget -- while streams join
and leave lanes at different steps.  The fake checks the lane bookkeeping (positions continue the lane's cache,
the fed-back frame is the lane's own last frame)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sesameai.serving import ContinuousBatcher, LaneGroups, Request

C, V = 32, 2051


def _frame(key: int, n: int, eos_at: int) -> torch.Tensor:
    if n == eos_at:
        return torch.zeros(C, dtype=torch.int32)
    return ((key * 131 + n * 7 + torch.arange(C)) % (V - 1) + 1).to(torch.int32)


def _expected(key: int, eos_at: int, budget: int) -> torch.Tensor:
    out = []
    for n in range(budget):
        f = _frame(key, n, eos_at)
        if bool((f == 0).all()):
            break
        out.append(f)
    return torch.stack(out) if out else torch.zeros(0, C, dtype=torch.int32)


class FakeModel(torch.nn.Module):
    def __init__(self, eos):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.config = type("Cfg", (), {"audio_num_codebooks": C})()
        self.eos = eos
        self.max_batch = 0
        self.calls = []
        self.pad_rows = 0

    def setup_caches(self, n):
        self.max_batch = n
        self.len = [0] * n
        self.key = [None] * n
        self.n = [0] * n
        self.last = [None] * n

    def reset_lane(self, lane):
        self.len[lane], self.key[lane], self.n[lane], self.last[lane] = 0, None, 0, None

    def lane_len(self, lane):
        return self.len[lane]

    def check_device_error(self):
        pass

    def clone_for_context(self):
        return FakeModel(self.eos)

    def generate_frame(self, tokens, mask, pos, temperature, topk, lanes=None):
        B, S, _ = tokens.shape
        lanes = list(range(B)) if lanes is None else list(lanes)
        assert len(set(lanes)) == B and all(0 <= l < self.max_batch for l in lanes)
        self.calls.append((B, S))
        out = torch.zeros(B, C, dtype=torch.int32)
        for b, l in enumerate(lanes):
            assert pos[b].tolist() == list(range(self.len[l], self.len[l] + S)), "positions must continue the lane"
            if S == 1 and self.key[l] is None:
                # an idle padding row (the batcher rounds the batch up to a captured graph size on free, rewound lanes)
                assert self.len[l] == 0 and not bool(mask[b, 0, C])
                self.len[l] += 1
                self.pad_rows += 1
                out[b] = 7
                continue
            if S > 1 or self.key[l] is None:
                assert self.len[l] == 0
                self.key[l] = int(tokens[b, 0, C])
                assert bool(mask[b, :, C].all())
            else:
                assert torch.equal(tokens[b, 0, :C].to(torch.int32), self.last[l]), "fed-back frame is not the lane's own"
                assert bool(mask[b, 0, :C].all()) and not bool(mask[b, 0, C])
            f = _frame(self.key[l], self.n[l], self.eos.get(self.key[l], -1))
            self.n[l] += 1
            self.len[l] += S
            self.last[l] = f
            out[b] = f
        return out


def _requests(n, budgets, prompt=lambda r: 2 + r % 3):
    reqs = []
    for r in range(n):
        S = prompt(r)
        tok = torch.zeros(S, C + 1, dtype=torch.long)
        tok[:, C] = r  # the fake keys a stream by its first text token
        msk = torch.zeros(S, C + 1, dtype=torch.bool)
        msk[:, C] = True
        reqs.append(Request(r, tok, msk, budgets[r]))
    return reqs


def test_join_leave_and_ragged_eos():
    n = 11
    eos = {0: 3, 1: 0, 4: 9, 7: 5}  # request 1 ends on its very first frame
    budgets = [6 + (r * 3) % 5 for r in range(n)]
    fm = FakeModel(eos)
    got = ContinuousBatcher(fm, 4, 0.9, 50).run(_requests(n, budgets))
    assert sorted(got) == list(range(n))
    for r in range(n):
        assert torch.equal(got[r], _expected(r, eos.get(r, -1), budgets[r])), r
    assert max(B for B, S in fm.calls if S == 1) == 4  # the lanes really ran batched
    assert {B for B, S in fm.calls if S == 1} <= {1, 2, 4}  # decode batches are padded to the captured sizes
    assert fm.pad_rows > 0
    assert sum(B for B, S in fm.calls if S > 1) == n   # every request joins through exactly one prefill row


def test_lane_groups_deal_round_robin():
    n = 9
    budgets = [4 + r % 3 for r in range(n)]
    fm = FakeModel({2: 1})
    got = LaneGroups(fm, 2, 2, 0.9, 50).run(_requests(n, budgets))
    for r in range(n):
        assert torch.equal(got[r], _expected(r, {2: 1}.get(r, -1), budgets[r])), r


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sesameai.parallel import serve_sharded

        budgets = [3 + r % 4 for r in range(n)]
        out = serve_sharded(FakeModel({3: 2}), _requests(n, budgets), rank, world, lanes=2, temperature=0.9, topk=50,
                            max_frames=16)
        if rank == 0:
            q.put(all(torch.equal(out[r], _expected(r, {3: 2}.get(r, -1), budgets[r])) for r in range(n)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_serve_sharded_two_gloo_ranks():
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


def test_throwaway_warm_up_server_leaves_the_next_server_untouched():
    """bench.py config 5 decodes two frames of every request on a throwaway server before the timed job (first-use
    costs of the process belong to the service's start-up): the pattern must not disturb the server that follows on
    the same model, and two-frame budgets must retire cleanly."""
    n = 13
    budgets = [5 + (r * 2) % 4 for r in range(n)]
    eos = {3: 2}
    fm = FakeModel(eos)
    reqs = _requests(n, budgets)
    warm = LaneGroups(fm, 1, 8, 0.9, 50, max_frames=16)
    got_w = warm.run([Request(q.rid, q.tokens, q.mask, 2) for q in reqs])
    assert sorted(got_w) == list(range(n)) and all(got_w[r].shape[0] <= 2 for r in range(n))
    del warm
    got = LaneGroups(fm, 1, 8, 0.9, 50, max_frames=16).run(reqs)
    for r in range(n):
        assert torch.equal(got[r], _expected(r, eos.get(r, -1), budgets[r])), r
