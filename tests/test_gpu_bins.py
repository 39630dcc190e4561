"""K8 (evx_bins.cuh): the slice's bin string built on the device.  The string, pushed through the
host coder, must give the same bits as the host's own binarisation of the table + records the same
frame produced -- and those are pinned against the oracle / reference elsewhere (test_gpu_api,
test_host_entropy).  Also: the stale-DC state across frames, the grow-and-emit-again path, R = 2/4/8."""
import numpy as np
import pytest

import oracleharness as O
from cairo_b200 import api, gpu, synth

pytestmark = pytest.mark.gpu


def _run(w, h, R, kinds, nframes, q, cap=None, linear=0):
    p = gpu.Pipeline(w, h, R, linear, 1)
    p.set_output(2)
    if cap:
        p.set_bins_capacity(cap)
    mbw, mbh = p.aw // 16, p.ah // 16
    wa, wb = api.SliceWriter(mbw, mbh, R), api.SliceWriter(mbw, mbh, R)
    for t in range(nframes):
        kind = kinds[t % len(kinds)]
        ft = 0 if t == 0 else 1
        if kind == "random":
            f = np.random.default_rng(w * 131 + h * 7 + t).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        elif kind == "grey":
            f = np.full((h, w, 3), 40 + 30 * t, np.uint8)
        else:
            f = synth.frame(w, h, t, 3, kind)
        p.encode_submit(f, ft, t, q)
        words, nbins, ncoded = p.encode_collect_bins()
        tbl, rec = p.encode_collect()
        assert ncoded == len(rec), (t, ncoded, len(rec))
        want, wbits = wa.serialize(tbl, rec)
        got, gbits = wb.serialize_bins(words, nbins)
        assert gbits == wbits and (got == want).all(), (w, h, R, t, kind, gbits, wbits)
    p.close()


@pytest.mark.parametrize("R", [2, 4, 8])
def test_bins_match_host_binarisation(R):
    _run(352, 288, R, ["moving", "static", "noise", "moving", "flat", "dark"], 8, 12)


def test_bins_stale_dc_across_frames():
    """Alternating still / moving content: copy blocks keep the DC of the frame that last coded them
    (serialize.cpp:59-72), which the device mirror must reproduce frame after frame."""
    _run(192, 160, 4, ["moving", "static", "static", "moving", "flat", "moving", "static", "noise", "static", "moving"], 12, 20)


def test_bins_ragged_and_tiny():
    for (w, h) in [(16, 16), (18, 34), (130, 18), (34, 130)]:
        _run(w, h, 4, ["random", "grey", "random", "grey"], 5, 8)


def test_bins_buffer_grows():
    """A bin buffer far too small: collect_bins must enlarge it and emit again, same bits."""
    _run(352, 288, 4, ["noise", "moving"], 3, 4, cap=4096)


def test_bins_linear_quant_high_quality():
    _run(320, 240, 4, ["noise", "moving", "dark"], 4, 2, linear=1)


def test_bins_1080p():
    _run(1920, 1080, 2, ["moving"], 3, 16)


def test_public_encoder_both_binarisations(monkeypatch):
    """evx1_encoder with the device bin string (default) and with EVX1_HOST_BINARISE=1: same stream."""
    w, h, n = 320, 240, 5
    frames = [synth.frame(w, h, t, 1, "moving") for t in range(n)]
    def encode_all():
        e = api.evx1_encoder(ref_count=4)
        e.set_quality(10)
        out = []
        for f in frames:
            data, bits = e.encode(f)
            out.append((bytes(data), bits))
        e.clear()
        return out
    a = encode_all()
    monkeypatch.setenv("EVX1_HOST_BINARISE", "1")
    b = encode_all()
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x == y


def test_two_frames_queued_on_the_device():
    """Bin-only output: a second frame may be queued behind the one in flight (evxgpu.h); the strings come back
    in order and equal the one-at-a-time run.  With string buffers below the worst case the second submit is refused."""
    from cairo_b200 import gpu
    w, h, q, n = 352, 288, 16, 6
    frames = [synth.frame(w, h, t, 5, "moving") for t in range(n)]
    a = gpu.Pipeline(w, h, 2, 0, 1)
    a.set_output(1)
    want = []
    for t in range(n):
        a.encode_submit(frames[t], 0 if t == 0 else 1, t, q)
        want.append(a.encode_collect_bins())
    b = gpu.Pipeline(w, h, 2, 0, 1)
    b.set_output(1)
    got = []
    b.encode_submit(frames[0], 0, 0, q)
    for t in range(1, n):
        b.encode_submit(frames[t], 1, t, q)
        with pytest.raises(RuntimeError):
            b.encode_submit(frames[t], 1, t + 1, q)          # a third frame
        got.append(b.encode_collect_bins())
    got.append(b.encode_collect_bins())
    with pytest.raises(RuntimeError):
        b.encode_collect_bins()
    for t in range(n):
        assert got[t][1] == want[t][1] and got[t][2] == want[t][2] and (got[t][0] == want[t][0]).all(), t
    c = gpu.Pipeline(w, h, 2, 0, 1)
    c.set_output(1)
    c.set_bins_capacity(1 << 12)                              # far below the worst case: one frame at a time
    c.encode_submit(frames[0], 0, 0, q)
    with pytest.raises(RuntimeError):
        c.encode_submit(frames[1], 1, 1, q)
    words, nbins, coded = c.encode_collect_bins()             # grows and emits again
    assert nbins == want[0][1] and (words == want[0][0]).all()
