"""Parity of the CUDA pixel pipeline (through the C-ABI, include/evxgpu.h) against the
C oracle and the committed golden vectors.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import goldenutil as G
import oracleharness as O
from cairo_b200 import synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all((np.asarray(x) == np.asarray(y)).all() for x, y in zip(a, b))


def _pipeline(*a, **k):
    from cairo_b200 import gpu
    return gpu.Pipeline(*a, **k)


def _run_encode(w, h, nf, R, lin, db, q, kind, seed=0, intra_every=0):
    from cairo_b200 import gpu
    p = _pipeline(w, h, R, lin, db)
    o = O.Oracle(w, h, R, lin, db)
    coef = [np.zeros((o.ah, o.aw), np.int16), np.zeros((o.ah // 2, o.aw // 2), np.int16), np.zeros((o.ah // 2, o.aw // 2), np.int16)]
    for t in range(nf):
        f = synth.frame(w, h, t, seed, kind)
        ft = 0 if (t == 0 or (intra_every and t % intra_every == 0)) else 1
        o.convert_in(f)
        o.encode_slice(ft, t, q)
        tbl, rec = p.encode(f, ft, t, q)
        assert _same(p.planes(0), o.planes(0)), f"yuv frame {t}"
        ot = o.block_table().copy()
        assert O.tables_equal(ot, tbl), f"block table frame {t}: {np.flatnonzero(ot['block_type'] != tbl['block_type'])[:8]}"
        gpu.records_to_planes(tbl, rec, coef, o.aw, o.ah)
        assert _same(coef, o.planes(1)), f"coefficients frame {t}"
        o.deblock(t)
        assert _same(p.planes(2, t % R), o.planes(2, t % R)), f"deblocked reconstruction frame {t}"
    p.close()


def test_convert_in_matches_oracle():
    for (w, h) in [(352, 288), (200, 120), (1920, 1080)]:
        p = _pipeline(w, h)
        o = O.Oracle(w, h)
        f = synth.frame(w, h, 2, 3, "noise")
        p.convert_in(f)
        o.convert_in(f)
        assert _same(p.planes(0), o.planes(0)), (w, h)
        p.close()


def test_inter_search_matches_oracle():
    w, h, q, R = 352, 288, 16, 4
    p = _pipeline(w, h, R)
    o = O.Oracle(w, h, R)
    for t in range(4):
        f = synth.frame(w, h, t, 1, "moving")
        o.convert_in(f); o.encode_slice(0 if t == 0 else 1, t, q); o.deblock(t)
    for s in range(R):
        for c in range(3):
            p.set_plane(2, s, c, o.plane(2, s, c))
    f = synth.frame(w, h, 4, 1, "moving")
    o.convert_in(f)
    p.convert_in(f)
    p.inter_search(4, q)
    for off in (1, 2, 3):
        d, sad = p.inter_result(off)
        for mb in range(p.nblocks):
            px, py = (mb % (o.aw // 16)) * 16, (mb // (o.aw // 16)) * 16
            od, osad = o.inter_prediction(4, q, px, py, off)
            assert osad == sad[mb] and O.tables_equal(np.array([od]), d[mb:mb + 1]), (off, mb, od, d[mb], osad, sad[mb])


@pytest.mark.parametrize("cfg", [(4, 0, 1), (2, 0, 1), (4, 1, 1), (4, 0, 0)])
def test_encode_cif_all_configs(cfg):
    _run_encode(352, 288, 5, *cfg, 16, "moving")


@pytest.mark.parametrize("kind", ["static", "flat", "noise", "dark"])
def test_encode_adversarial(kind):
    _run_encode(176, 144, 5, 4, 0, 1, 16, kind, seed=1)


@pytest.mark.parametrize("q", [1, 7, 8, 24, 31])
def test_encode_quality_sweep(q):
    _run_encode(176, 144, 4, 4, 0, 1, q, "moving", seed=2)


def test_encode_unaligned_and_periodic_intra():
    _run_encode(200, 120, 6, 4, 0, 1, 16, "moving", seed=3, intra_every=3)
    _run_encode(200, 120, 4, 2, 0, 1, 12, "noise", seed=4)


def test_deblock_kernel_random():
    rng = np.random.default_rng(5)
    w, h = 96, 80
    p = _pipeline(w, h)
    o = O.Oracle(w, h)
    for trial in range(12):
        span = [4, 16, 64, 300][trial % 4]
        for comp in range(3):
            pa = o.plane(2, 0, comp)
            pa[...] = rng.integers(-span, span, size=pa.shape, dtype=np.int16) + 128
            p.set_plane(2, 0, comp, pa)
        t = o.block_table()
        t["block_type"] = rng.integers(0, 8, size=t.shape[0])
        t["q_index"] = rng.integers(1, 32, size=t.shape[0])
        p.set_block_table(t.copy())
        o.deblock(0)
        p.deblock(0)
        assert _same(p.planes(2, 0), o.planes(2, 0)), trial


@pytest.mark.parametrize("name", G.names())
def test_golden_encode_and_decode(name):
    from cairo_b200 import gpu
    g = G.Golden(name)
    enc = _pipeline(g.w, g.h, g.R, g.linear, g.deblocking)
    dec = _pipeline(g.w, g.h, g.R, g.linear, g.deblocking)
    aw, ah = enc.aw, enc.ah
    coef = [np.zeros((ah, aw), np.int16), np.zeros((ah // 2, aw // 2), np.int16), np.zeros((ah // 2, aw // 2), np.int16)]
    for t in range(g.frames):
        ft = 0 if g.is_intra(t) else 1
        tbl, rec = enc.encode(g.rgb(t), ft, t, g.q)
        assert O.tables_equal(g.table(t), tbl), t
        gpu.records_to_planes(tbl, rec, coef, aw, ah)
        assert _same(coef, g.planes(t, "coef")), t
        assert _same(enc.planes(2, t % g.R), g.planes(t, "deblocked")), t
        # decoder: feed the golden table + golden coefficients
        gt = g.table(t)
        grec = gpu.planes_to_records(gt, g.planes(t, "coef"), aw)
        rgb = dec.decode(gt, grec, ft, t)
        assert _same(dec.planes(2, t % g.R), g.planes(t, "deblocked")), t
        assert (rgb == g.decoded_rgb(t)).all(), t


def test_1080p_encode_matches_oracle_two_frames():
    """BASELINE config size; the oracle needs ~2 s per 1080p P-frame, so two frames only."""
    _run_encode(1920, 1080, 2, 2, 0, 1, 16, "moving", seed=5)


def test_decoder_dependency_tracking_on_adversarial_tables():
    """Random block tables: intra-motion blocks pointing anywhere in the frame under construction
    (already-decoded blocks AND not-yet-decoded ones, whose stale ring contents must be read),
    random sub-pel, random coefficients.  The GPU decoder must equal the oracle's raster-order
    decode_slice.  Sources never overlap the block's own position (no valid stream does that)."""
    from cairo_b200 import gpu
    rng = np.random.default_rng(11)
    w, h, R = 128, 96, 4
    mbw, mbh = w // 16, h // 16
    p = gpu.Pipeline(w, h, R, 0, 1)
    o = O.Oracle(w, h, R, 0, 1)
    for t in range(6):
        tbl = np.zeros(mbw * mbh, dtype=gpu.BLOCK_DESC_DTYPE)
        for mb in range(mbw * mbh):
            bx, by = mb % mbw, mb // mbw
            ty = int(rng.choice([1, 3, 7, 3, 7, 0, 2, 4, 6]))       # intra-motion types over-represented
            tbl["block_type"][mb] = ty
            tbl["prediction_target"][mb] = rng.integers(1, R) if not (ty & 1) else 0
            if ty & 2:
                while True:
                    sx, sy = int(rng.integers(1, w - 17)), int(rng.integers(1, h - 17))
                    if abs(sx - bx * 16) >= 18 or abs(sy - by * 16) >= 18:
                        break
                tbl["motion_x"][mb], tbl["motion_y"][mb] = sx - bx * 16, sy - by * 16
                tbl["sp_pred"][mb] = rng.integers(0, 2)
                tbl["sp_amount"][mb] = rng.integers(0, 2)
                tbl["sp_index"][mb] = rng.integers(0, 8)
            tbl["q_index"][mb] = rng.integers(1, 32)
        n = int(((tbl["block_type"] & 4) == 0).sum())
        rec = rng.integers(-40, 40, size=(n, 384)).astype(np.int16)
        rec[rng.random(rec.shape) < 0.85] = 0
        # oracle: table + coefficient planes, then the reference's raster-order decode
        o.block_table()[...] = tbl
        coef = [o.plane(1, 0, c) for c in range(3)]
        gpu.records_to_planes(tbl, rec, coef, o.aw, o.ah)
        o.decode_slice(0 if t == 0 else 1, t)
        o.deblock(t)
        rgb = p.decode(tbl, rec, 0 if t == 0 else 1, t)
        assert _same(p.planes(2, t % R), o.planes(2, t % R)), t
        assert (rgb == o.convert_out(t)).all(), t


@pytest.mark.parametrize("size", [(2, 2), (16, 16), (18, 34), (48, 16), (16, 48), (30, 30), (130, 18), (34, 130)])
def test_tiny_and_ragged_frames(size):
    """Smallest and ragged frames: one macroblock, one row, one column, widths that are not
    multiples of 8 (byte-path colour conversion) -- encode parity and decode round trip."""
    from cairo_b200 import gpu
    w, h = size
    R, q = 4, 12
    p = gpu.Pipeline(w, h, R, 0, 1)
    dec = gpu.Pipeline(w, h, R, 0, 1)
    o = O.Oracle(w, h, R, 0, 1)
    rng = np.random.default_rng(w * 1000 + h)
    for t in range(4):
        f = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) if t % 2 else np.full((h, w, 3), 40 + 50 * t, np.uint8)
        ft = 0 if t == 0 else 1
        o.convert_in(f)
        o.encode_slice(ft, t, q)
        tbl, rec = p.encode(f, ft, t, q)
        assert _same(p.planes(0), o.planes(0)), (size, t)
        assert O.tables_equal(o.block_table().copy(), tbl), (size, t)
        o.deblock(t)
        assert _same(p.planes(2, t % R), o.planes(2, t % R)), (size, t)
        rgb = dec.decode(tbl, rec, ft, t)
        assert (rgb == o.convert_out(t)).all(), (size, t)


def test_create_rejects_bad_geometry():
    from cairo_b200 import gpu
    for w, h in [(0, 16), (16, 0), (15, 16), (16, 15), (4096, 4096)]:      # odd sizes; > 65535 macroblocks
        with pytest.raises(RuntimeError):
            gpu.Pipeline(w, h)
    with pytest.raises(RuntimeError):
        gpu.Pipeline(64, 64, ref_count=1)        # a ring of one slot has no past frame
    with pytest.raises(RuntimeError):
        gpu.Pipeline(64, 64, ref_count=9)


def test_largest_frame_geometry_runs():
    """4096x4080 = 65280 macroblocks, just under the uint16 block_count limit (serialize.cpp:321)."""
    from cairo_b200 import gpu
    w, h = 4096, 4080
    p = gpu.Pipeline(w, h, 2, 0, 1)
    assert p.nblocks == 65280
    f = synth.frame(w, h, 0, 0, "moving")
    tbl, rec = p.encode(f, 0, 0, 16)
    tbl2, rec2 = p.encode(f, 1, 1, 16)
    assert ((tbl2["block_type"] & 4) != 0).mean() > 0.5      # a repeated frame is mostly copy blocks
