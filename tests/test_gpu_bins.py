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
