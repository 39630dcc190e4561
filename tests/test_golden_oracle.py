"""The C oracle against the committed golden vectors (reference output, see
tests/golden/make_golden.py).  Runs anywhere -- /root/reference is not needed."""
import pytest

import goldenutil as G
import oracleharness as O


def _same(a, b):
    return all((x == y).all() for x, y in zip(a, b))


@pytest.mark.parametrize("name", G.names())
def test_oracle_reproduces_golden(name):
    g = G.Golden(name)
    enc = O.Oracle(g.w, g.h, g.R, g.linear, g.deblocking)
    dec = O.Oracle(g.w, g.h, g.R, g.linear, g.deblocking)
    for t in range(g.frames):
        ft = 0 if g.is_intra(t) else 1
        enc.convert_in(g.rgb(t))
        assert _same(enc.planes(0), g.planes(t, "src"))
        enc.encode_slice(ft, t, g.q)
        assert O.tables_equal(g.table(t), enc.block_table().copy())
        assert _same(enc.planes(1), g.planes(t, "coef"))
        assert _same(enc.planes(2, t % g.R), g.planes(t, "recon"))
        d, b = enc.serialize()
        gd, gb = g.slice_bits(t)
        assert O.bits_equal(d, b, gd, gb)
        enc.deblock(t)
        assert _same(enc.planes(2, t % g.R), g.planes(t, "deblocked"))
        # decoder side from the golden slice
        dec.unserialize(gd, gb)
        dec.decode_slice(ft, t)
        dec.deblock(t, tiled=True)
        assert _same(dec.planes(2, t % g.R), g.planes(t, "deblocked"))
        assert (dec.convert_out(t) == g.decoded_rgb(t)).all()
