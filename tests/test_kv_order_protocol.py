"""Exhaustive interleaving check of the KV-ordering hand-shake between a CTA's consumer warps and its producer warp
(sesameai-tts_b200/csrc/mega.cuh: kv_post / kv_producer_poll / the producer's drain loop) on a small model:
three CTAs, three codebook steps.  Each CTA has

  * a consumer program (thread NCT-1 of mega::k_frame_mega):  per step  post(release, tag) ; post(acquire, tag) ;
    wait(ack == tag|acquire) ; ... ; finally post(exit).  A post first waits until the previous request was acknowledged;
  * a producer program: read the request word; if it changed: exit -> retire; release -> done[cta] = tag;
    acquire -> block until every CTA's done word carries the tag; then acknowledge.

The data dependencies of the frame are modelled by one rule: a CTA starts the work of step i + 1 (and therefore can
reach that step's release) only after EVERY CTA has passed its wait of step i (every gate/up phase needs every CTA).
Checked over ALL interleavings: no deadlock, every request served exactly once and in order, an acquire is only
acknowledged after every CTA released that tag, no done word is overwritten while a CTA still polls for the old tag,
both programs terminate.  This is a model of the protocol, not of the CUDA code; the placement of the flags in the real
phase table is checked by tests/test_phase_table.py, the kernel itself by the GPU parity / stress tests."""
from collections import deque

NCTA, STEPS = 3, 3
REL, ACQ, EXIT = 1, 2, 3


def _consumer_program():
    prog = []
    prev = 0
    for i in range(1, STEPS + 1):
        tag = i << 4
        prog.append(("post", tag | REL, prev))
        prog.append(("post", tag | ACQ, tag | REL))
        prog.append(("wait", tag | ACQ, i))
        prev = tag | ACQ
    prog.append(("exit",))
    return prog


PROG = _consumer_program()


def _initial():
    # per CTA: consumer pc, producer state (0 idle / 1 serving), producer's last seen word, request, ack, done word,
    # number of waits passed; plus the served log per CTA (as a tuple)
    cta = (0, 0, 0, 0, 0, 0, 0, ())
    return tuple(cta for _ in range(NCTA))


def _moves(state):
    out = []
    for c, (pc, pst, last, req, ack, done, passed, log) in enumerate(state):
        # ---- consumer ----
        if pc < len(PROG):
            ins = PROG[pc]
            if ins[0] == "post":
                word, prev = ins[1], ins[2]
                step = word >> 4
                deps_ok = True
                if word & 3 == REL and step > 1:  # the work of step `step` needs every CTA past its wait of step - 1
                    deps_ok = all(s[6] >= step - 1 for s in state)
                if ack == prev and deps_ok:
                    out.append((c, (pc + 1, pst, last, word, ack, done, passed, log)))
            elif ins[0] == "wait":
                if ack == ins[1]:
                    out.append((c, (pc + 1, pst, last, req, ack, done, ins[2], log)))
            else:  # exit request: the last acquire was acknowledged (the wait before it), nothing to wait for
                out.append((c, (pc + 1, pst, last, EXIT, ack, done, passed, log)))
        # ---- producer ----
        if pst == 0:
            if req != last:
                if req == EXIT:
                    out.append((c, (pc, 2, req, req, ack, done, passed, log)))  # retired
                else:
                    out.append((c, (pc, 1, req, req, ack, done, passed, log)))
        elif pst == 1:
            tag, typ = last & ~3, last & 3
            if typ == REL:
                # nobody may still be polling for the word this store overwrites
                for o, s in enumerate(state):
                    if s[1] == 1 and s[2] & 3 == ACQ and (s[2] & ~3) == done and done != tag:
                        raise AssertionError(f"CTA {c} overwrites done word {done:#x} while CTA {o} polls for it")
                out.append((c, (pc, 0, last, req, last, tag, passed, log + (last,))))
            else:
                if all(s[5] == tag for s in state):  # every CTA's done word carries the tag
                    out.append((c, (pc, 0, last, req, last, done, passed, log + (last,))))
    return out


def test_request_acknowledge_protocol_all_interleavings():
    want_log = tuple(w for i in range(1, STEPS + 1) for w in ((i << 4) | REL, (i << 4) | ACQ))
    seen = {_initial()}
    todo = deque(seen)
    finals = 0
    while todo:
        st = todo.popleft()
        mv = _moves(st)
        if not mv:
            # terminal: everything ran to completion
            for pc, pst, last, req, ack, done, passed, log in st:
                assert pc == len(PROG), ("deadlock: consumer stuck", st)
                assert pst == 2, ("deadlock: producer did not retire", st)
                assert log == want_log, ("requests lost, duplicated or reordered", log)
                assert passed == STEPS
            finals += 1
            continue
        for c, new in mv:
            nxt = st[:c] + (new,) + st[c + 1:]
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    assert finals >= 1
    assert len(seen) > 500  # the search really covered interleavings
