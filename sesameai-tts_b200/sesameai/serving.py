"""Continuous batching on top of ``Model.generate_frame``.

The reference serves ONE stream per ``Generator.generate`` call: prompt, then one ``generate_frame`` per 80 ms
until the all-zero EOS frame or the frame budget (``sesameai/generator.py:283-294``); its batch semantics are
lock-step (one shared cache position, ``torch.all(sample == 0)`` over the whole batch, ``models.py:160``).
Here every request owns a KV-cache *lane* with its own length: a queued request joins by prefilling a free lane
(a batch-1 call), all active lanes advance together in one batched decode call per frame step, and a stream
leaves at ITS EOS frame or budget, freeing the lane for the next request.  Per stream the frames are exactly
what the reference loop would have produced for that request alone (same prompt layout, same EOS rule, same
feedback of the sampled frame as the next input row).

``LaneGroups`` runs several such batchers on separate CUDA streams and separate decode contexts over the same
parameter tensors: a decode step is a chain of ~770 short dependent launches, so independent groups overlap in
the launch gaps and latency bubbles of each other.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Deque, Dict, Iterable, List, Optional, Sequence

import torch

from .models import Model


@dataclass
class Request:
    """One utterance: prompt frames [S, 33] int64 + mask [S, 33] bool (``Generator._tokenize_*`` layout)."""

    rid: int
    tokens: torch.Tensor
    mask: torch.Tensor
    max_frames: int


@dataclass
class _Stream:
    req: Request
    lane: int
    n_frames: int = 0  # frames kept so far (rows of the batcher's frame buffer)
    calls: int = 0     # generate_frame calls spent (the reference's loop counter)


class ContinuousBatcher:
    def __init__(self, model: Model, max_lanes: int, temperature: float, topk: int, *, setup: bool = True,
                 stream: Optional[torch.cuda.Stream] = None, max_frames: int = 2048):
        self.model = model
        self.max_lanes = int(max_lanes)
        self.temperature, self.topk = float(temperature), int(topk)
        if setup:
            model.setup_caches(self.max_lanes)
        self.device = next(model.parameters()).device
        self.stream = stream
        C = model.config.audio_num_codebooks
        self._C = C
        self._last = torch.zeros(self.max_lanes, C, dtype=torch.int32, device=self.device)  # newest frame per lane
        # frames of the running streams, on the device: one scatter per round, one slice + copy per finished stream
        self._max_frames = int(max_frames)
        self._frames = torch.zeros(self.max_lanes, self._max_frames, C, dtype=torch.int32, device=self.device)
        # per-lane bookkeeping mirrored on the device, so that a round builds its inputs with a handful of tensor ops
        # instead of Python loops and host -> device copies (the GPU idles while the host prepares a round)
        self._len_dev = torch.zeros(self.max_lanes, dtype=torch.int64, device=self.device)  # cache length (= model.lane_len)
        self._nf_dev = torch.zeros(self.max_lanes, dtype=torch.int64, device=self.device)   # frames kept
        self._io: Dict[int, tuple] = {}  # padded batch size -> (tokens, mask) buffers
        self._free: List[int] = list(range(self.max_lanes - 1, -1, -1))
        self._active: Dict[int, _Stream] = {}
        self._pending: Deque[Request] = deque()
        self._done: Dict[int, torch.Tensor] = {}
        self._rows: List[int] = []
        self._rows_dev: Optional[torch.Tensor] = None
        self._inflight = None
        self.steps = 0          # batched decode calls issued
        self.row_steps = 0      # sum of their batch sizes (= frames computed by decode calls)
        self.prefills = 0       # prefill calls (requests of equal prompt length join in one batched call)

    # -- queue -------------------------------------------------------------------------------------------
    def submit(self, req: Request) -> None:
        if req.max_frames > self._max_frames:
            raise ValueError("request asks for more frames than the batcher was built for")
        self._pending.append(req)

    @property
    def idle(self) -> bool:
        return not self._pending and not self._active and self._inflight is None

    def results(self) -> Dict[int, torch.Tensor]:
        return self._done

    # -- one scheduling round: begin() launches asynchronously, end() looks at the EOS flags -------------------
    def _ctx(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else _Null()

    def _lane_len(self, lane: int) -> int:
        return self.model.lane_len(lane)

    def begin(self) -> None:
        assert self._inflight is None
        m = self.model
        with self._ctx():
            # advance: one batched decode step for every stream that already holds a frame to feed back
            rows = sorted(self._active)
            if rows:
                if rows != self._rows:
                    self._rows = rows
                    self._rows_dev = torch.tensor(rows, dtype=torch.long, device=self.device)
                B = len(rows)
                # The library replays one captured CUDA graph per batch size: as streams leave one by one the
                # batch is padded to the next bucket with idle rows on free lanes (rewound before and after), so a
                # ragged pool needs a dozen graphs instead of one per distinct size.
                pad = self._free[: max(0, min(_bucket(B), self.max_lanes) - B)]
                for l in pad:
                    m.reset_lane(l)
                Bp = B + len(pad)
                if Bp not in self._io:
                    msk = torch.ones(Bp, 1, self._C + 1, dtype=torch.bool, device=self.device)
                    msk[:, :, -1] = False
                    self._io[Bp] = (torch.zeros(Bp, 1, self._C + 1, dtype=torch.int64, device=self.device), msk)
                tok, msk = self._io[Bp]
                tok.zero_()
                tok[:B, 0, : self._C] = self._last.index_select(0, self._rows_dev)
                pos = torch.zeros(Bp, 1, dtype=torch.int64, device=self.device)  # idle rows sit on rewound lanes: position 0
                pos[:B, 0] = self._len_dev.index_select(0, self._rows_dev)
                out = m.generate_frame(tok, msk, pos, self.temperature, self.topk, lanes=rows + pad)[:B]
                for l in pad:
                    m.reset_lane(l)
                self._len_dev.index_add_(0, self._rows_dev, torch.ones_like(self._rows_dev))
                self._last.index_copy_(0, self._rows_dev, out)
                for l in rows:
                    self._active[l].calls += 1
                self.steps += 1
                self.row_steps += B
            # join: queued requests take the free lanes; those with the same prompt length are prefilled by ONE
            # batched call (the prompt's last row samples each stream's first frame)
            fresh: List[_Stream] = []
            while self._pending and self._free:
                fresh.append(_Stream(self._pending.popleft(), self._free.pop()))
            by_len: Dict[int, List[_Stream]] = {}
            for st in fresh:
                by_len.setdefault(int(st.req.tokens.shape[0]), []).append(st)
            for S, group in by_len.items():
                lanes = [st.lane for st in group]
                for l in lanes:
                    m.reset_lane(l)
                tok = torch.stack([st.req.tokens for st in group]).to(self.device, torch.int64)
                msk = torch.stack([st.req.mask for st in group]).to(self.device, torch.bool)
                pos = torch.arange(S, device=self.device).unsqueeze(0).repeat(len(group), 1)
                s = m.generate_frame(tok, msk, pos, self.temperature, self.topk, lanes=lanes)
                lanes_t = torch.tensor(lanes, dtype=torch.long, device=self.device)
                self._last.index_copy_(0, lanes_t, s)
                self._len_dev[lanes_t] = S
                self._nf_dev[lanes_t] = 0
                for st in group:
                    st.calls = 1
                    self._active[st.lane] = st
                self.prefills += 1
            lanes_now = sorted(self._active)
            if not lanes_now:
                return
            idx = self._rows_dev if lanes_now == self._rows else torch.tensor(lanes_now, dtype=torch.long, device=self.device)
            newest = self._last.index_select(0, idx)                      # [n, C] frames sampled this round
            eos = (newest == 0).all(dim=1)                                # reference generator.py:285, per stream
            # store the frame of every stream at its own frame index (an EOS frame lands one past the end and is never read)
            at = self._nf_dev.index_select(0, idx).clamp_(max=self._max_frames - 1)
            self._frames[idx, at] = newest
            self._nf_dev.index_add_(0, idx, (~eos).to(torch.int64))  # (the host mirror advances in end())
            eos_h = eos.to("cpu", non_blocking=True)
            ev = None
            if self.device.type == "cuda":
                ev = torch.cuda.Event()
                ev.record()
            self._inflight = (lanes_now, eos_h, ev)

    def end(self) -> None:
        if self._inflight is None:
            return
        lanes_now, eos, ev = self._inflight
        self._inflight = None
        if ev is not None:
            ev.synchronize()
        self.model.check_device_error()
        eos = eos.tolist()
        for i, lane in enumerate(lanes_now):
            st = self._active[lane]
            finished = eos[i]
            if not finished:
                st.n_frames += 1
                finished = st.calls >= st.req.max_frames
            if finished:
                self._done[st.req.rid] = self._frames[lane, : st.n_frames].to("cpu", copy=True)
                del self._active[lane]
                self._free.append(lane)

    def run(self, requests: Iterable[Request]) -> Dict[int, torch.Tensor]:
        for r in requests:
            self.submit(r)
        while not self.idle:
            self.begin()
            self.end()
        return self._done


def _bucket(b: int) -> int:
    """Batch sizes the decode graphs are captured for: 1, 2, 4, 8, multiples of 8 up to 64, then multiples of 32."""
    if b <= 8:
        return 1 if b <= 1 else 2 if b <= 2 else 4 if b <= 4 else 8
    if b <= 64:
        return (b + 7) // 8 * 8
    return (b + 31) // 32 * 32


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class LaneGroups:
    """``groups`` independent ContinuousBatchers (own decode context + CUDA stream each, shared parameters);
    requests are dealt round-robin.  All groups launch their round before any of them waits for its EOS flags."""

    def __init__(self, model: Model, groups: int, lanes_per_group: int, temperature: float, topk: int, max_frames: int = 2048):
        self.batchers: List[ContinuousBatcher] = []
        dev = next(model.parameters()).device
        for g in range(groups):
            mg = model if g == 0 else model.clone_for_context()
            st = torch.cuda.Stream(device=dev) if groups > 1 and dev.type == "cuda" else None
            if st is not None:
                st.wait_stream(torch.cuda.current_stream(dev))
            with (torch.cuda.stream(st) if st is not None else _Null()):
                self.batchers.append(ContinuousBatcher(mg, lanes_per_group, temperature, topk, stream=st, max_frames=max_frames))

    def run(self, requests: Sequence[Request]) -> Dict[int, torch.Tensor]:
        for i, r in enumerate(requests):
            self.batchers[i % len(self.batchers)].submit(r)
        while not all(b.idle for b in self.batchers):
            for b in self.batchers:
                if not b.idle:
                    b.begin()
            for b in self.batchers:
                b.end()
        out: Dict[int, torch.Tensor] = {}
        for b in self.batchers:
            out.update(b.results())
        return out

    @property
    def steps(self) -> int:
        return sum(b.steps for b in self.batchers)

    @property
    def row_steps(self) -> int:
        return sum(b.row_steps for b in self.batchers)
