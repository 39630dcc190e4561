"""The shipping path pinned to the reference at BASELINE.json's full sizes.

The shipping path is evx1_encoder::submit/collect with the device-side bin string, several frame slots and
consecutive frames overlapping on the device.  Here it is compared, frame by frame and byte by byte, with the
UNMODIFIED reference (oracle/_ref/libevxref_<variant>.so: evx1_encoder::encode, evx1enc.cpp:92-156, and
evx1_decoder::decode, evx1dec.cpp:87-123) -- or, on a box without oracle/_ref, with the C port of the oracle,
which tests/test_oracle_vs_ref.py pins to the same reference stage by stage.

  configs[1]  1080p, ring of 2                      12 frames
  configs[2]  1080p, ring of 4, MPEG + adaptive QP   12 frames, and the linear (uniform) quantiser variant
  configs[3]  3840x2160, insert_intra every 3rd frame, decode round trip
H7 masks (stream byte 7, the unused bits of the last byte) as in tests/test_gpu_api.py."""
import numpy as np
import pytest

import oracleharness as O
import refharness as R
from cairo_b200 import synth

pytestmark = pytest.mark.gpu


def _streams_equal(a, abits, b, bbits, first):
    if abits != bbits:
        return False
    ua = np.unpackbits(np.asarray(a, np.uint8), bitorder="little")[:abits].copy()
    ub = np.unpackbits(np.asarray(b, np.uint8), bitorder="little")[:bbits].copy()
    if first:
        ua[56:64] = 0
        ub[56:64] = 0
    return bool((ua == ub).all())


class _Reference:
    """Frame-by-frame reference streams and decoded pictures: the compiled reference when it is here, else the C port."""

    def __init__(self, variant, w, h, ring, linear, q):
        self.w, self.h, self.q, self.t = w, h, q, 0
        self.use_ref = R.available(variant)
        if self.use_ref:
            self.enc, self.dec = R.RefEncoder(variant), R.RefDecoder(variant)
            self.enc.set_quality(q)
        else:
            self.o = O.Oracle(w, h, ring, linear, 1)
        self.intra_next = True

    def insert_intra(self):
        self.intra_next = True
        if self.use_ref:
            self.enc.insert_intra()

    def frame(self, rgb):
        """-> (compare(data, bits) -> bool, decoded RGB of the frame)"""
        t, intra = self.t, self.intra_next
        self.t += 1
        self.intra_next = False
        if self.use_ref:
            rd, rb = self.enc.encode(rgb)
            pic = self.dec.decode(rd, rb, self.w, self.h)
            return (lambda d, b: _streams_equal(d, b, rd, rb, t == 0)), pic
        self.o.convert_in(rgb)
        self.o.encode_slice(0 if intra else 1, t, self.q)
        od, ob = self.o.serialize()
        self.o.deblock(t)
        pic = self.o.convert_out(t)
        skip = (24 if t == 0 else 10) * 8

        def cmp(d, b):
            got = np.packbits(np.unpackbits(np.asarray(d, np.uint8), bitorder="little")[skip:b], bitorder="little")
            return O.bits_equal(got, b - skip, od, ob)
        return cmp, pic


def _run_pipelined(w, h, ring, linear, variant, q, n, seed, look, intra_every=0, kind="moving"):
    from cairo_b200 import api
    frames = [synth.frame(w, h, t, seed, kind) for t in range(n)]
    enc = api.evx1_encoder(ref_count=ring, linear_quant=linear)
    enc.set_quality(q)
    dec = api.evx1_decoder(linear_quant=linear)
    ref = _Reference(variant, w, h, ring, linear, q)
    out = []

    def take():
        d, b = enc.collect()
        out.append((d.copy(), b))

    for t in range(n):
        if intra_every and t and t % intra_every == 0:
            enc.insert_intra()
        enc.submit(frames[t])
        if t >= look:
            take()
    while len(out) < n:
        take()
    for t in range(n):
        if intra_every and t and t % intra_every == 0:
            ref.insert_intra()
        cmp, pic = ref.frame(frames[t])
        d, b = out[t]
        assert cmp(d, b), (variant, t, b)
        rgb = dec.decode(d, b, w, h)
        assert (rgb == pic).all(), (variant, t)


@pytest.mark.parametrize("variant,ring,linear", [("r2", 2, 0), ("r4", 4, 0), ("r4_linear", 4, 1)])
def test_1080p_pipelined_stream_equals_reference(variant, ring, linear):
    """configs[1] / configs[2]: 12 frames of 1080p through submit/collect (frames overlapping on the device, five
    frames of lookahead), every frame's bytes and decoded picture against the reference."""
    _run_pipelined(1920, 1080, ring, linear, variant, 16, 12, 5, look=5)


def test_1080p_pipelined_noise_and_dark_content_equals_reference():
    """The adversarial generators at full size through the shipping path: full-range noise (INTRA_DEFAULT, large
    coefficients, long bin strings) and dark frames (the sad < 8192 tie rule)."""
    _run_pipelined(1920, 1080, 2, 0, "r2", 24, 4, 7, look=3, kind="noise")
    _run_pipelined(1920, 1080, 2, 0, "r2", 8, 5, 9, look=4, kind="dark")


def test_4k_periodic_intra_stream_equals_reference():
    """configs[3]: 3840x2160, insert_intra every 3rd frame, pipelined; bytes and decode round trip against the reference."""
    _run_pipelined(3840, 2160, 2, 0, "r2", 16, 5, 1, look=3, intra_every=3)
