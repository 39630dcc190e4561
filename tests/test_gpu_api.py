"""The reference-facing API (evx1_encoder::encode / evx1_decoder::decode / set_quality /
insert_intra, evx1.h:66-113) end to end on the GPU: byte-identical streams and decoded frames
against the golden vectors, the C oracle, and (when oracle/_ref is present) the reference itself."""
import numpy as np
import pytest

import goldenutil as G
import oracleharness as O
import refharness as R
from cairo_b200 import synth

pytestmark = pytest.mark.gpu


def _streams_equal(a, abits, b, bbits, first):
    """Bit-length comparison; stream byte 7 (padding inside evx_header) is masked."""
    if abits != bbits:
        return False
    ua = np.unpackbits(np.asarray(a, np.uint8), bitorder="little")[:abits].copy()
    ub = np.unpackbits(np.asarray(b, np.uint8), bitorder="little")[:bbits].copy()
    if first:
        ua[56:64] = 0
        ub[56:64] = 0
    return bool((ua == ub).all())


@pytest.mark.parametrize("name", G.names())
def test_public_api_matches_golden_streams(name):
    from cairo_b200 import api
    g = G.Golden(name)
    enc = api.evx1_encoder(ref_count=g.R, linear_quant=g.linear, deblocking=g.deblocking)
    enc.set_quality(g.q)
    dec = api.evx1_decoder(linear_quant=g.linear, deblocking=g.deblocking)
    for t in range(g.frames):
        if t and g.is_intra(t):
            enc.insert_intra()
        data, bits = enc.encode(g.rgb(t))
        gd, gb = g.stream(t)
        assert _streams_equal(data, bits, gd, gb, t == 0), (name, t, bits, gb)
        rgb = dec.decode(gd, gb, g.w, g.h)
        assert (rgb == g.decoded_rgb(t)).all(), (name, t)


def test_cif_30_frames_vs_reference_or_oracle():
    """BASELINE configs[0]: CIF 352x288, 30 frames, quality 16, encode+decode."""
    from cairo_b200 import api
    w, h, q, n = 352, 288, 16, 30
    enc = api.evx1_encoder()
    enc.set_quality(q)
    dec = api.evx1_decoder()
    use_ref = R.available("r4")
    if use_ref:
        renc, rdec = R.RefEncoder("r4"), R.RefDecoder("r4")
        renc.set_quality(q)
    else:
        o = O.Oracle(w, h, 4, 0, 1)
    for t in range(n):
        f = synth.frame(w, h, t, 0, "moving")
        data, bits = enc.encode(f)
        rgb = dec.decode(data, bits, w, h)
        if use_ref:
            rd, rb = renc.encode(f)
            assert _streams_equal(data, bits, rd, rb, t == 0), t
            assert (rgb == rdec.decode(rd, rb, w, h)).all(), t
        else:
            o.convert_in(f)
            o.encode_slice(0 if t == 0 else 1, t, q)
            od, ob = o.serialize()
            o.deblock(t)
            skip = (24 if t == 0 else 10) * 8
            assert O.bits_equal(np.packbits(np.unpackbits(data, bitorder="little")[skip:bits], bitorder="little"), bits - skip, od, ob), t
            assert (rgb == o.convert_out(t)).all(), t


def test_set_quality_and_clear_semantics():
    from cairo_b200 import api
    w, h = 96, 80
    enc = api.evx1_encoder()
    f = synth.frame(w, h, 0, 0, "moving")
    d0, b0 = enc.encode(f)                       # default quality 8 (config.h:43)
    hdr = np.frombuffer(bytes(d0[:24]), dtype=np.uint8)
    assert bytes(hdr[:4]) == b"EVX1" and hdr[6] == 4                         # magic, ref_count
    assert int(hdr[10]) | (int(hdr[11]) << 8) == w and int(hdr[12]) | (int(hdr[13]) << 8) == h
    assert int(hdr[22]) == 8                                                 # frame.quality, low byte
    enc.set_quality(200)                          # clipped to 31 (evx1enc.cpp:53-64)
    d1, b1 = enc.encode(f)
    assert int(d1[8]) == 31 and int(d1[0]) == 1 and int(d1[4]) == 1          # quality, type inter, index 1
    enc.clear()                                   # next frame restarts the stream with a header
    d2, b2 = enc.encode(f)
    assert bytes(d2[:4]) == b"EVX1" and _streams_equal(d0, b0, d2, b2, True)


def test_decoder_rejects_out_of_order_frames():
    from cairo_b200 import api
    w, h = 96, 80
    enc = api.evx1_encoder()
    dec = api.evx1_decoder()
    a = enc.encode(synth.frame(w, h, 0, 0, "moving"))
    a = (a[0].copy(), a[1])
    b = enc.encode(synth.frame(w, h, 1, 0, "moving"))
    b = (b[0].copy(), b[1])
    c = enc.encode(synth.frame(w, h, 2, 0, "moving"))
    c = (c[0].copy(), c[1])
    dec.decode(a[0], a[1], w, h)
    with pytest.raises(RuntimeError):
        dec.decode(c[0], c[1], w, h)              # index 2 while 1 is expected (evx1dec.cpp:77-80)
    dec.decode(b[0], b[1], w, h)


def test_1080p_round_trip_property():
    """Full-size property: the encoder's in-loop reconstruction equals what the decoder rebuilds
    from the stream (closed loop), for a P-frame chain at BASELINE's 1080p size."""
    from cairo_b200 import api
    w, h = 1920, 1080
    enc = api.evx1_encoder(ref_count=2)
    enc.set_quality(16)
    dec = api.evx1_decoder()
    prev = None
    for t in range(4):
        f = synth.frame(w, h, t, 0, "moving")
        data, bits = enc.encode(f)
        rgb = dec.decode(data, bits, w, h)
        err = np.abs(rgb.astype(np.int32) - f.astype(np.int32)).mean()
        assert err < 6.0, (t, err)                # lossy but close at quality 16
        prev = rgb


@pytest.mark.parametrize("name", G.names())
def test_pipelined_submit_collect_matches_golden_streams(name):
    """submit(n+1) before collect(n): the frames leave one call later and are the same bytes, including
    set_quality / insert_intra issued between the two halves (they belong to the next submit)."""
    from cairo_b200 import api
    g = G.Golden(name)
    enc = api.evx1_encoder(ref_count=g.R, linear_quant=g.linear, deblocking=g.deblocking)
    enc.set_quality(g.q)
    frames = [np.ascontiguousarray(g.rgb(t)) for t in range(g.frames)]
    enc.submit(frames[0])
    for t in range(1, g.frames + 1):
        if t < g.frames:
            if g.is_intra(t):
                enc.insert_intra()
            enc.submit(frames[t])
        data, bits = enc.collect()
        gd, gb = g.stream(t - 1)
        assert _streams_equal(data, bits, gd, gb, t == 1), (name, t - 1, bits, gb)


def test_pipelined_1080p_equals_synchronous_and_state_rules():
    from cairo_b200 import api
    w, h, n = 1920, 1080, 12
    frames = [synth.frame(w, h, t, 3, "moving") for t in range(n)]
    a = api.evx1_encoder(ref_count=2)
    a.set_quality(16)
    sync = []
    for t in range(n):
        d, b = a.encode(frames[t])
        sync.append((d.copy(), b))
    p = api.evx1_encoder(ref_count=2)
    p.set_quality(16)
    with pytest.raises(RuntimeError):
        p.collect()                               # nothing submitted: EVX_ERROR_NOT_READY
    p.submit(frames[0])
    with pytest.raises(RuntimeError):
        p.encode(frames[1])                       # encode() with a frame uncollected
    held = 1
    while held < n:                               # frames on the device (two or three) + six retired ones being coded
        try:
            p.submit(frames[held])
        except RuntimeError:                      # one uncollected frame too many: EVX_ERROR_NOT_READY
            break
        held += 1
    assert held >= 8, held                        # (frame slots on the device + retired frames being coded)
    out = []
    for t in range(held, n):                      # that much lookahead from here on
        d, b = p.collect()
        out.append((d.copy(), b))
        p.submit(frames[t])
    while len(out) < n:
        d, b = p.collect()
        out.append((d.copy(), b))
    with pytest.raises(RuntimeError):
        p.collect()
    d, b = p.encode(frames[0])                    # back to the synchronous call once drained
    assert b > 0
    for t in range(n):
        assert out[t][1] == sync[t][1] and (out[t][0] == sync[t][0]).all(), t


@pytest.mark.parametrize("env", [{"EVXGPU_FRAME_SLOTS": "3"}, {"EVXGPU_FRAME_SLOTS": "2"}, {"EVXGPU_FRAME_SLOTS": "8"}, {"EVXGPU_FRAME_OVERLAP": "0"}])
def test_pipelined_session_frame_slots(env, monkeypatch):
    """evx1_encoder::submit/collect with two, three and eight frame slots and without the frame pipeline: the same bytes as encode()."""
    from cairo_b200 import api
    w, h, n = 640, 368, 14
    frames = [synth.frame(w, h, t, 11, "moving") for t in range(n)]
    a = api.evx1_encoder(ref_count=2)
    a.set_quality(12)
    sync = []
    for t in range(n):
        d, b = a.encode(frames[t])
        sync.append((d.copy(), b))
    del a
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    p = api.evx1_encoder(ref_count=2)
    p.set_quality(12)
    out = []
    look = 9
    for t in range(n):
        p.submit(frames[t])
        if t >= look:
            d, b = p.collect()
            out.append((d.copy(), b))
    while len(out) < n:
        d, b = p.collect()
        out.append((d.copy(), b))
    for t in range(n):
        assert out[t][1] == sync[t][1] and (out[t][0] == sync[t][0]).all(), (env, t)


def test_pipelined_decode_equals_synchronous():
    """evx1_decoder::submit(n+1) before collect(n): same pictures as decode(), one call later; state rules."""
    from cairo_b200 import api
    w, h, n = 352, 288, 16
    enc = api.evx1_encoder()
    enc.set_quality(16)
    streams = []
    for t in range(n):
        if t == 5:
            enc.insert_intra()
        d, b = enc.encode(synth.frame(w, h, t, 1, "moving"))
        streams.append((d.copy(), b))
    ref = api.evx1_decoder()
    want = [ref.decode(d, b, w, h).copy() for d, b in streams]
    dec = api.evx1_decoder()
    with pytest.raises(RuntimeError):
        dec.collect(w, h)                          # nothing submitted
    dec.submit(*streams[0])
    with pytest.raises(RuntimeError):
        dec.decode(streams[1][0], streams[1][1], w, h)     # decode() with a frame uncollected
    held = 1
    for t in range(1, 12):
        dec.submit(*streams[t])                    # twelve frames may be uncollected (their slices parse concurrently)
        held += 1
    with pytest.raises(RuntimeError):
        dec.submit(*streams[held])                 # a thirteenth
    got = [dec.collect(w, h).copy()]
    for t in range(held, n):
        dec.submit(*streams[t])
        got.append(dec.collect(w, h).copy())
    for _ in range(held - 1):
        got.append(dec.collect(w, h).copy())
    with pytest.raises(RuntimeError):
        dec.collect(w, h)
    for t in range(n):
        assert (got[t] == want[t]).all(), t


def test_peek_views_match_reference():
    """evx1_encoder::peek (evx1enc.cpp:170-305): the six implemented debug views against the golden pictures
    of the reference (tests/golden/make_golden_peek.py) and, when oracle/_ref is present, the reference itself."""
    import os
    from cairo_b200 import api
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peek", "peek_176x144.npz"))
    w, h, q, n = int(g["w"]), int(g["h"]), int(g["q"]), int(g["frames"])
    enc = api.evx1_encoder()
    assert (enc.peek(enc.PEEK_SOURCE, w, h) == 0).all()          # not initialised yet: success, nothing written
    enc.set_quality(q)
    renc = R.RefEncoder("r4") if R.available("r4") and hasattr(R.lib("r4"), "evxref_encoder_peek") else None
    if renc:
        renc.set_quality(q)
    for t in range(n):
        f = synth.frame(w, h, t, 0, "moving")
        enc.encode(f)
        if renc:
            renc.encode(f)
    views = {"source": 0, "block_table": 2, "quant_table": 3, "spmp_table": 4, "block_variance": 5, "destination": 6}
    for name, state in views.items():
        got = enc.peek(state, w, h)
        assert (got == g[name]).all(), name
        if renc:
            assert (got == renc.peek(state, w, h)).all(), name
    with pytest.raises(RuntimeError):
        enc.peek(enc.PEEK_PREDICTION, w, h)                       # no case in the reference's switch: EVX_ERROR_NOTIMPL
    enc.submit(synth.frame(w, h, n, 0, "moving"))
    with pytest.raises(RuntimeError):
        enc.peek(enc.PEEK_SOURCE, w, h)                           # a pipelined frame is uncollected
    enc.collect()
    assert enc.peek(enc.PEEK_DESTINATION, w, h).any()
