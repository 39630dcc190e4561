"""The two re-orderings the GPU relies on, proven on the CPU oracle first:

 * macroblocks in wavefront order (step = bx + 3*by) give the raster-order result;
 * deblocking as independent 8x8 tiles centred on the grid crossings gives the result of the
   reference's in-place raster sweep (deblock.cpp:201-254).
"""
import numpy as np
import pytest

import oracleharness as O
from cairo_b200 import synth


def _same(a, b):
    return all((x == y).all() for x, y in zip(a, b))


@pytest.mark.parametrize("kind,q,R", [("moving", 16, 4), ("noise", 31, 4), ("dark", 8, 2), ("static", 16, 4)])
@pytest.mark.parametrize("reverse", [0, 1])
def test_wavefront_order_equals_raster(kind, q, R, reverse):
    w, h = 208, 160
    a = O.Oracle(w, h, R, 0, 1)
    b = O.Oracle(w, h, R, 0, 1)
    for t in range(6):
        f = synth.frame(w, h, t, 21, kind)
        ft = 0 if t == 0 else 1
        a.convert_in(f); b.convert_in(f)
        a.encode_slice(ft, t, q)
        b.encode_slice(ft, t, q, wavefront=reverse)
        assert O.tables_equal(a.block_table().copy(), b.block_table().copy()), t
        assert _same(a.planes(1), b.planes(1)), t
        assert _same(a.planes(2, t % R), b.planes(2, t % R)), t
        a.deblock(t); b.deblock(t)


@pytest.mark.parametrize("kind,q", [("moving", 16), ("noise", 31), ("noise", 8), ("dark", 24), ("moving", 7)])
def test_tiled_deblock_equals_raster_sweep(kind, q):
    w, h = 208, 168   # 168 -> aligned 176
    a = O.Oracle(w, h, 4, 0, 1)
    b = O.Oracle(w, h, 4, 0, 1)
    changed = 0
    for t in range(5):
        f = synth.frame(w, h, t, 22, kind)
        ft = 0 if t == 0 else 1
        for o in (a, b):
            o.convert_in(f)
            o.encode_slice(ft, t, q)
        before = a.planes(2, t % 4)
        a.deblock(t, tiled=False)
        b.deblock(t, tiled=True)
        assert _same(a.planes(2, t % 4), b.planes(2, t % 4)), t
        changed += sum(int((x != y).sum()) for x, y in zip(before, a.planes(2, t % 4)))
    if q >= 8:
        assert changed > 0, "deblocking never fired; the test would be vacuous"


def test_tiled_deblock_on_random_planes_and_tables():
    """Adversarial: random reconstruction, random copy flags and qp -- every strength/qp combination."""
    rng = np.random.default_rng(5)
    w, h = 96, 80
    a = O.Oracle(w, h, 4, 0, 1)
    b = O.Oracle(w, h, 4, 0, 1)
    for trial in range(20):
        span = [4, 16, 64, 300][trial % 4]
        for comp in range(3):
            pa = a.plane(2, 0, comp)
            pa[...] = rng.integers(-span, span, size=pa.shape, dtype=np.int16) + 128
            b.plane(2, 0, comp)[...] = pa
        ta, tb = a.block_table(), b.block_table()
        ta["block_type"] = rng.integers(0, 8, size=ta.shape[0])
        ta["q_index"] = rng.integers(1, 32, size=ta.shape[0])
        tb[...] = ta
        a.deblock(0, tiled=False)
        b.deblock(0, tiled=True)
        assert _same(a.planes(2, 0), b.planes(2, 0)), trial
