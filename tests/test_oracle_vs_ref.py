"""Pins the plain-C oracle (oracle/evx_oracle.c) to the UNMODIFIED reference compiled
by oracle/Makefile (oracle/_ref/libevxref_<variant>.so): every stage, bit for bit."""
import numpy as np
import pytest

import oracleharness as O
import refharness as R
from cairo_b200 import synth

VARIANTS = {  # name -> (ring slots, linear quant, deblocking)
    "r4": (4, 0, 1),
    "r2": (2, 0, 1),
    "r4_linear": (4, 1, 1),
    "r4_nodeblock": (4, 0, 0),
}


def _need(variant):
    if not R.available(variant):
        pytest.skip(f"oracle/_ref/libevxref_{variant}.so not built (no /root/reference here)")


def _same(a, b):
    return all((x == y).all() for x, y in zip(a, b))


def _run(w, h, nf, variant, q, kind, seed=0, intra_every=0):
    _need(variant)
    Rn, lin, db = VARIANTS[variant]
    st = R.RefStage(w, h, variant)
    o = O.Oracle(w, h, Rn, lin, db)
    for t in range(nf):
        f = synth.frame(w, h, t, seed, kind)
        ft = 0 if (t == 0 or (intra_every and t % intra_every == 0)) else 1
        st.set_frame(ft, t, q)
        st.convert_in(f)
        o.convert_in(f)
        assert _same(st.planes(0), o.planes(0)), f"yuv frame {t}"
        st.encode_slice()
        o.encode_slice(ft, t, q)
        assert O.tables_equal(st.block_table(), o.block_table().copy()), f"block table frame {t}"
        assert _same(st.planes(1), o.planes(1)), f"coefficients frame {t}"
        assert _same(st.planes(2, t % Rn), o.planes(2, t % Rn)), f"reconstruction frame {t}"
        d1, b1 = st.serialize()
        d2, b2 = o.serialize()
        assert O.bits_equal(d1, b1, d2, b2), f"slice bits frame {t}"
        st.deblock()
        o.deblock(t)
        assert _same(st.planes(2, t % Rn), o.planes(2, t % Rn)), f"deblocked frame {t}"
        assert (st.convert_out() == o.convert_out(t)).all(), f"rgb frame {t}"


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_cif_moving_all_variants(variant):
    _run(352, 288, 5, variant, 16, "moving")


@pytest.mark.parametrize("kind", ["static", "flat", "noise", "dark"])
def test_adversarial_content(kind):
    _run(176, 144, 5, "r4", 16, kind, seed=1)


@pytest.mark.parametrize("q", [1, 7, 8, 24, 31])
def test_quality_sweep(q):
    _run(176, 144, 4, "r4", q, "moving", seed=2)


def test_unaligned_size_and_periodic_intra():
    # 200x120: neither dimension a multiple of 16 (padding rows/cols stay 0)
    _run(200, 120, 6, "r4", 16, "moving", seed=3, intra_every=3)
    _run(200, 120, 4, "r2", 12, "noise", seed=4)


def test_linear_quant_noise():
    _run(176, 144, 3, "r4_linear", 20, "noise", seed=5)


def test_decoder_path_matches_reference():
    """unserialize -> decode_slice -> deblock -> RGB against the reference decoder stages."""
    _need("r4")
    w, h, q = 176, 144, 16
    enc = R.RefStage(w, h, "r4")
    dref = R.RefStage(w, h, "r4")
    o = O.Oracle(w, h, 4, 0, 1)
    for t in range(5):
        f = synth.frame(w, h, t, 7, "moving")
        ft = 0 if t == 0 else 1
        enc.set_frame(ft, t, q)
        enc.convert_in(f)
        enc.encode_slice()
        data, bits = enc.serialize()
        enc.deblock()
        dref.set_frame(ft, t, q)
        dref.unserialize(data, bits)
        dref.decode_slice()
        dref.deblock()
        o.unserialize(data, bits)
        assert O.tables_equal(dref.block_table(), o.block_table().copy())
        assert _same(dref.planes(0), o.planes(1))      # decoder coefficients live in input_cache
        o.decode_slice(ft, t)
        o.deblock(t)
        assert _same(dref.planes(2, t % 4), o.planes(2, t % 4))
        assert _same(enc.planes(2, t % 4), o.planes(2, t % 4))   # encoder/decoder closed loop
        assert (dref.convert_out() == o.convert_out(t)).all()


def test_public_api_stream_equals_staged_stream():
    """evx1_encoder::encode = 14-byte header (first frame) + 10-byte frame desc + slice (evx1enc.cpp:92-156)."""
    _need("r4")
    w, h, q = 176, 144, 16
    enc = R.RefEncoder("r4")
    enc.set_quality(q)
    dec = R.RefDecoder("r4")
    o = O.Oracle(w, h, 4, 0, 1)
    for t in range(4):
        f = synth.frame(w, h, t, 9, "moving")
        data, bits = enc.encode(f)
        o.convert_in(f)
        o.encode_slice(0 if t == 0 else 1, t, q)
        d2, b2 = o.serialize()
        o.deblock(t)
        skip = (24 if t == 0 else 10) * 8
        a = np.unpackbits(data, bitorder="little")[skip:bits]
        b = np.unpackbits(d2, bitorder="little")[:b2]
        assert a.size == b.size and (a == b).all()
        assert (dec.decode(data, bits, w, h) == o.convert_out(t)).all()


def test_single_searches_match_reference():
    _need("r4")
    w, h, q = 176, 144, 16
    st = R.RefStage(w, h, "r4")
    o = O.Oracle(w, h, 4, 0, 1)
    for t in range(3):
        f = synth.frame(w, h, t, 11, "moving")
        ft = 0 if t == 0 else 1
        for x in (st, ):
            x.set_frame(ft, t, q); x.convert_in(f); x.encode_slice(); x.deblock()
        o.convert_in(f); o.encode_slice(ft, t, q); o.deblock(t)
    f = synth.frame(w, h, 3, 11, "moving")
    st.set_frame(1, 3, q); st.convert_in(f); o.convert_in(f)
    for py in range(0, 144, 16):
        for px in range(0, 176, 16):
            for off in (1, 2, 3):
                d1, s1 = st.inter_prediction(px, py, off)
                d2, s2 = o.inter_prediction(3, q, px, py, off)
                assert s1 == s2 and O.tables_equal(np.array([d1]), np.array([d2])), (px, py, off)
            d1, s1 = st.intra_prediction(px, py)
            d2, s2 = o.intra_prediction(3, q, px, py)
            assert s1 == s2 and O.tables_equal(np.array([d1]), np.array([d2])), (px, py)
