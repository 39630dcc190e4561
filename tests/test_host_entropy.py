"""Host entropy stage (cairo_b200/csrc/host/entropy.cpp) against the golden slices made by the
reference: the writer must reproduce the reference's bits from the reference's block table and
coefficients, the reader must invert them -- including the stale-DC semantics across frames
(SURVEY H4).  No GPU needed."""
import os

import numpy as np
import pytest

import goldenutil as G
import oracleharness as O
from cairo_b200 import api, gpu


@pytest.mark.parametrize("name", G.names())
def test_slice_writer_reproduces_reference_bits(name):
    g = G.Golden(name)
    aw = (g.w + 15) // 16 * 16
    ah = (g.h + 15) // 16 * 16
    wr = api.SliceWriter(aw // 16, ah // 16, g.R)
    for t in range(g.frames):
        tbl = g.table(t)
        rec = gpu.planes_to_records(tbl, g.planes(t, "coef"), aw)
        d, b = wr.serialize(tbl, rec)
        gd, gb = g.slice_bits(t)
        assert O.bits_equal(d, b, gd, gb), (name, t, b, gb)


@pytest.mark.parametrize("name", G.names())
def test_slice_reader_inverts_reference_bits(name):
    g = G.Golden(name)
    aw = (g.w + 15) // 16 * 16
    ah = (g.h + 15) // 16 * 16
    rd = api.SliceReader(aw // 16, ah // 16, g.R)
    coef = [np.zeros((ah, aw), np.int16), np.zeros((ah // 2, aw // 2), np.int16), np.zeros((ah // 2, aw // 2), np.int16)]
    for t in range(g.frames):
        gd, gb = g.slice_bits(t)
        tbl, rec = rd.unserialize(gd, gb)
        assert O.tables_equal(g.table(t), tbl, check_variance=False), (name, t)   # variance is not on the wire
        gpu.records_to_planes(tbl, rec, coef, aw, ah)
        assert all((a == b).all() for a, b in zip(coef, g.planes(t, "coef"))), (name, t)


def test_writer_reader_round_trip_random_tables():
    rng = np.random.default_rng(3)
    mbw, mbh, R = 7, 5, 4
    wr, rd = api.SliceWriter(mbw, mbh, R), api.SliceReader(mbw, mbh, R)
    for trial in range(8):
        tbl = np.zeros(mbw * mbh, dtype=gpu.BLOCK_DESC_DTYPE)
        tbl["block_type"] = rng.integers(0, 8, size=tbl.shape[0])
        tbl["prediction_target"] = rng.integers(0, R, size=tbl.shape[0])
        tbl["motion_x"] = rng.integers(-40, 40, size=tbl.shape[0])
        tbl["motion_y"] = rng.integers(-40, 40, size=tbl.shape[0])
        tbl["sp_pred"] = rng.integers(0, 2, size=tbl.shape[0])
        tbl["sp_amount"] = rng.integers(0, 2, size=tbl.shape[0])
        tbl["sp_index"] = rng.integers(0, 8, size=tbl.shape[0])
        tbl["q_index"] = rng.integers(1, 32, size=tbl.shape[0])
        n = int(((tbl["block_type"] & 4) == 0).sum())
        rec = rng.integers(-300, 300, size=(n, 384)).astype(np.int16)
        rec[rng.random(rec.shape) < 0.8] = 0
        if n:
            rec[0, 5] = 32767 if trial % 2 else -32767      # long escape codes
        d, b = wr.serialize(tbl, rec)
        t2, r2 = rd.unserialize(d, b)
        assert O.tables_equal(tbl, t2, check_variance=False)
        assert (r2 == rec).all()


def test_empty_frame_all_copy_blocks():
    mbw, mbh, R = 4, 3, 2
    wr, rd = api.SliceWriter(mbw, mbh, R), api.SliceReader(mbw, mbh, R)
    tbl = np.zeros(mbw * mbh, dtype=gpu.BLOCK_DESC_DTYPE)
    tbl["block_type"] = 4          # INTER_COPY everywhere: no vectors, no q, no residuals
    tbl["prediction_target"] = 1
    d, b = wr.serialize(tbl, np.zeros((0, 384), np.int16))
    t2, r2 = rd.unserialize(d, b)
    assert (t2["block_type"] == 4).all() and r2.shape[0] == 0


def test_fast_coder_against_plain_coder(tmp_path):
    """profiles/abac_bench.cpp `check`: the production arithmetic-coder loop (reciprocal split point,
    batched E1/E2, byte-drained accumulator, >8M-bin tail) against a bit-at-a-time coder written from the
    algorithm, over skewed/bursty bin strings; and the split-point identity for every tot < 2^23."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abac_bench")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "cairo_b200", "csrc", "host"),
                           "-o", exe, os.path.join(root, "profiles", "abac_bench.cpp"), "-lpthread"])
    out = subprocess.run([exe, "check"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all ok" in out.stdout, out.stdout[-2000:]


def test_reader_matches_oracle_on_truncated_and_corrupt_slices():
    """The fast arithmetic decoder against the oracle's bit-at-a-time restatement of the reference
    (abac.cpp:123-154, 226-279, 398-420) where their semantics are most particular: slices cut short
    (past the end the last bit read in the call repeats), bit flips, and pure noise (values outside
    [low, high], collapsed intervals, absurd run lengths).  Both sides start from the same state."""
    g = G.Golden(G.names()[0])
    aw, ah = (g.w + 15) // 16 * 16, (g.h + 15) // 16 * 16
    rng = np.random.default_rng(11)
    cases = []
    for t in range(min(g.frames, 3)):
        gd, gb = g.slice_bits(t)
        gd = np.array(gd, dtype=np.uint8)
        cases.append((gd, gb))
        for cut in (gb - 1, gb - 9, gb - 17, gb // 2, 40, 17, 16, 3, 0):
            if 0 <= cut < gb:
                cases.append((gd, cut))
        for _ in range(4):
            bad = gd.copy()
            for pos in rng.integers(0, gb, size=3):
                bad[pos >> 3] ^= np.uint8(1 << (pos & 7))
            cases.append((bad, gb))
    for nbits in (64, 1000, 20000):
        cases.append((rng.integers(0, 256, size=nbits // 8 + 8, dtype=np.uint8), nbits))
    for data, nbits in cases:
        o = O.Oracle(g.w, g.h, g.R, 0, 1)
        rd = api.SliceReader(aw // 16, ah // 16, g.R)
        o.unserialize(data, nbits)
        tbl, rec = rd.unserialize(data, nbits)
        assert O.tables_equal(o.block_table().copy(), tbl, check_variance=False), nbits
        want = gpu.planes_to_records(tbl, o.planes(1), aw)
        assert want.shape == rec.shape and (want == rec).all(), nbits


@pytest.mark.parametrize("name", G.names())
def test_slices_parse_in_any_order_on_any_thread(name):
    """The decoder parses the slices of consecutive frames concurrently (slice_reader::parse keeps no state) and
    merges them in frame order (apply): parsing every frame of a golden sequence up front, in reverse order and
    on several threads, then applying in order, must give what frame-by-frame unserialize gives."""
    import threading
    g = G.Golden(name)
    mbw, mbh = (g.w + 15) // 16, (g.h + 15) // 16
    seq = api.SliceReader(mbw, mbh, g.R)
    slices = [g.slice_bits(t) for t in range(g.frames)]          # (the .npz reader is not thread-safe)
    want = [seq.unserialize(*slices[t]) for t in range(g.frames)]
    rd = api.SliceReader(mbw, mbh, g.R)
    parsed = [None] * g.frames

    def work(ts):
        for t in ts:
            parsed[t] = rd.parse(*slices[t])

    order = list(range(g.frames))[::-1]
    threads = [threading.Thread(target=work, args=(order[k::3],)) for k in range(3)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for t in range(g.frames):
        tbl, rec = rd.apply(parsed[t])
        rd.free_parsed(parsed[t])
        for field in tbl.dtype.names:             # field by field: the descriptor has a padding byte
            assert (tbl[field] == want[t][0][field]).all(), (name, t, field)
        assert rec.shape == want[t][1].shape and (rec == want[t][1]).all(), (name, t)
