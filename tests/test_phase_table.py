"""Static check of the megakernel's tagged hand-off (mega.cuh) on the host-built phase table:
  * every tagged vector a phase consumes names, as its source, the phase that wrote that vector last;
  * no vector is rewritten sooner than two phases after it was written (a consumer of phase p's words
    may still be reading them while phase p+1 runs; seeing a word of phase p+1 proves phase p is over);
  * the consumer of a vector comes after its producer.
  * the depth decoder's PLAIN cache rows (written by a QKV epilogue or the sampling phase's table gather, read by
    other CTAs in later codebook steps of the same launch): between every writer and every reader there is a release
    (all CTAs, after the writer phase) and an acquire (all CTAs, on that release's done words, a whole phase before
    the reader's prefetch) -- mega::kv_step_sync.
No GPU needed: csm_debug_phase_table builds the table against an imaginary workspace."""
import ctypes as C

import pytest

from sesameai import _native

GEMV, EMBED, ATTN, SAMPLE = 0, 1, 2, 3
PLAIN, RESID, SWIGLU, ROPE_KV = 0, 1, 2, 3


def _cfg(tiny):
    cfg = _native.Config()
    if tiny:
        bb, dec = (2, 256, 4, 2, 512), (2, 256, 2, 1, 512)
    else:
        bb, dec = (16, 2048, 32, 8, 8192), (4, 1024, 8, 2, 8192)
    for dst, v in ((cfg.backbone, bb), (cfg.decoder, dec)):
        dst.layers, dst.dim, dst.heads, dst.kv_heads, dst.ff = v
    cfg.text_vocab, cfg.audio_vocab, cfg.codebooks, cfg.max_seq_len, cfg.norm_eps = (1000 if tiny else 128256), 2051, 32, 2048, 1e-5
    return cfg


def _table(cfg, ncta, qkv):
    buf = (_native.PhaseInfo * 700)()
    n = _native.lib().csm_debug_phase_table(C.byref(cfg), ncta, qkv, buf, 700)
    assert n > 0, _native.lib().csm_last_error()
    return [buf[i] for i in range(n)]


@pytest.mark.parametrize("tiny", [True, False])
@pytest.mark.parametrize("ncta", [148, 132, 16])
@pytest.mark.parametrize("qkv", [0, 1])
def test_tagged_hand_off_invariants(tiny, ncta, qkv):
    cfg = _cfg(tiny)
    ph = _table(cfg, ncta, qkv)
    layers_bb, layers_dec = cfg.backbone.layers, cfg.decoder.layers
    assert len(ph) == 1 + 5 * layers_bb + 2 + 31 * (4 * layers_dec + 2) - (30 if qkv else 0)
    kv_words = lambda s: 2 * s.kv_heads * (s.dim // s.heads)
    writer = {}  # (vector base offset + row offset in words) -> phase that wrote it last

    def consume(p, base, row_words, row, src):
        key = base + 4 * row * row_words
        assert key in writer, (p, "consumes a vector nobody wrote")
        assert writer[key] == src, (p, "source", src, "but last writer", writer[key])
        assert src < p

    def produce(p, base, row_words, row, in_place_only_reader=False):
        key = base + 4 * row * row_words
        if key in writer:
            # an in-place residual update may follow its source directly when the epilogue thread that rewrites a
            # word is the only reader of the old one (the [q;k;v] table removes the QKV phase that used to sit between)
            need = 1 if in_place_only_reader else 2
            assert p - writer[key] >= need, (p, "rewrites a vector written by phase", writer[key])
        writer[key] = p

    for p, x in enumerate(ph):
        is_dec = p > 1 + 5 * layers_bb + 1
        st = cfg.decoder if is_dec else cfg.backbone
        if x.type == EMBED:
            produce(p, x.t_out, 0, 0)
        elif x.type == ATTN:
            consume(p, x.t_q, 0, 0, x.q_src)
            consume(p, x.t_kv, 0, 0, x.q_src)
            produce(p, x.t_out, 0, 0)
        elif x.type == SAMPLE:
            consume(p, x.t_logits, 0, 0, x.logits_src)
            if x.t_next:
                produce(p, x.t_next, 0, 0)
            if x.has_qkv_table:
                produce(p, x.t_q, 0, 0)
                produce(p, x.t_kv, 0, 0)
        else:
            assert x.type == GEMV
            if x.attn_prologue:
                for n in range(x.nb):
                    consume(p, x.t_q, st.dim, n, x.q_src)
                    consume(p, x.t_kv, kv_words(st), n, x.q_src)
            else:
                for n in range(x.nb):
                    consume(p, x.t_x, x.ldx, n, x.x_src[n])
            if x.epi == RESID:
                for n in range(x.nb):
                    consume(p, x.t_out, x.ldo, n, x.resid_src[n])
                    reads_it_as_x = (not x.attn_prologue) and x.t_x == x.t_out
                    produce(p, x.t_out, x.ldo, n, in_place_only_reader=not reads_it_as_x)
            elif x.epi == ROPE_KV:
                for n in range(x.nb):
                    produce(p, x.t_q, st.dim, n)
                    produce(p, x.t_kv, kv_words(st), n)
            else:
                for n in range(x.nb):
                    produce(p, x.t_out, x.ldo, n)
                if x.t_out2:
                    produce(p, x.t_out2, 0, 0)
    # every GEMV phase fits the kernel's per-CTA limits on a full GPU (setup_mega falls back to the per-op path otherwise)
    for x in ph:
        if x.type == GEMV:
            assert x.R in (8, 16) and x.K % 256 == 0
            if ncta >= 132:
                assert -(-x.G // ncta) <= 8


@pytest.mark.parametrize("tiny", [True, False])
@pytest.mark.parametrize("ncta", [148, 16])
@pytest.mark.parametrize("qkv", [0, 1])
def test_plain_kv_rows_are_released_and_acquired(tiny, ncta, qkv):
    """Flags sit on the NEXT phase: kv_sync on phase f runs after the barrier that ends phase f - 1, in every CTA.
    A reader phase p prefetches its cached rows after the barrier that ends phase p - 1."""
    cfg = _cfg(tiny)
    ph = _table(cfg, ncta, qkv)
    first_dec = 1 + 5 * cfg.backbone.layers + 2
    releases = {}           # done_src -> flagged phase
    acquired_release = -1   # flagged phase of the release that the latest acquire observed
    acquire_at = -1
    waited_at = -1          # phase flagged with the wait for the latest acquire's acknowledgement
    writers_before_step = -1  # last phase that wrote plain rows the current step reads from the cache
    last_writer = -1
    readers = 0
    for p, x in enumerate(ph):
        if x.kv_sync == 1:
            assert x.done_src not in releases, "done-word tags must be unique within a frame"
            assert 0 <= x.done_src < p
            releases[x.done_src] = p
        elif x.kv_sync == 2:
            assert x.done_src in releases and releases[x.done_src] < p, (p, "acquire without an earlier release")
            acquired_release, acquire_at = releases[x.done_src], p
        elif x.kv_sync == 3:
            # producer-warp variant: every consumer thread waits here for the acknowledgement of that acquire, right
            # before its own prefetch of phase p's cached rows
            assert acquire_at >= 0 and ph[acquire_at].done_src == x.done_src and acquire_at < p
            waited_at = p
        else:
            assert x.kv_sync == 0
        if p < first_dec:
            continue
        if x.type == SAMPLE:
            # rows written up to here (this phase's own gather excluded: its row travels as tagged words in the next
            # step and is read from the cache only in the step after) are what the next step reads from the cache
            writers_before_step = last_writer
            if x.has_qkv_table:
                last_writer = p
        elif x.type == GEMV and x.epi == ROPE_KV and x.pos_mode == 0:
            last_writer = p
        if x.type == GEMV and x.attn_prologue and x.pos_mode == 0 and x.pos0 > 0:
            readers += 1
            assert writers_before_step >= 0
            # release runs after phase (flag - 1): every writer phase must be <= flag - 1
            assert acquired_release > writers_before_step, (p, "reads rows of phase", writers_before_step, "released at", acquired_release)
            # the acquiring warp and the prefetching threads are ordered by the barrier ending phase acquire_at
            assert acquire_at < p, (p, "acquire at", acquire_at)
            # producer-warp variant: the wait for that acquire's acknowledgement sits at or before the reader
            assert acquire_at < waited_at <= p, (p, "wait at", waited_at)
    assert readers == (cfg.codebooks - 2) * cfg.decoder.layers
