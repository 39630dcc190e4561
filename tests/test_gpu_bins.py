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
    w, h, q, n = 352, 288, 16, 14
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
    cap = b.encode_capacity()                                 # the handle's frame slots (ten by default)
    assert 2 <= cap <= 16
    for t in range(cap - 1):
        b.encode_submit(frames[t], 0 if t == 0 else 1, t, q)
    for t in range(cap - 1, n):
        b.encode_submit(frames[t], 1, t, q)
        with pytest.raises(RuntimeError):
            b.encode_submit(frames[t], 1, t + 1, q)          # one frame more than the handle holds
        got.append(b.encode_collect_bins())
    for _ in range(cap - 1):
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


def _same_bins(a, b):
    """(words, nbins, coded) triples: equal bin strings (bits past nbins in the last word are not part of the string)."""
    if a[1] != b[1] or a[2] != b[2]:
        return False
    n = a[1]
    full, rest = n // 64, n % 64
    if not (a[0][:full] == b[0][:full]).all():
        return False
    if rest:
        mask = np.uint64((1 << rest) - 1)
        return bool((a[0][full] & mask) == (b[0][full] & mask))
    return True


def _bins_sequence(env, w, h, R, frames, types, qualities):
    """The bin strings of a sequence through the C-ABI with two frames queued, under the given environment."""
    import os
    from cairo_b200 import gpu
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        p = gpu.Pipeline(w, h, R, 0, 1)
        p.set_output(1)
        out = []
        ahead = p.encode_capacity() - 1           # frames kept queued behind the one being collected
        inflight = 0
        for t in range(len(frames)):
            while True:
                try:
                    p.encode_submit(frames[t], types[t], t, qualities[t])
                    inflight += 1
                    break
                except RuntimeError as ex:        # status 8: every frame slot is taken
                    assert "status 8" in str(ex) and inflight > 0, ex
                    out.append(p.encode_collect_bins())
                    inflight -= 1
            if inflight > ahead:
                out.append(p.encode_collect_bins())
                inflight -= 1
        while inflight:
            out.append(p.encode_collect_bins())
            inflight -= 1
        rec = [a.copy() for a in p.planes(2, (len(frames) - 1) % R)]
        p.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out, rec


@pytest.mark.parametrize("w,h,R", [(1920, 1080, 2), (640, 368, 4), (176, 144, 2), (112, 32, 2), (64, 48, 3)])
def test_overlapped_frames_equal_serial_frames(w, h, R):
    """The frame pipeline (consecutive frames of a stream concurrently on the device, each one launch of the frame kernel
    -- search role, wavefront rows, deblocking behind the wavefront -- gated macroblock by macroblock on its predecessor):
    bin strings and the last reconstruction equal the frame-after-frame run of the stand-alone kernels -- with intra
    frames in the middle, quality changes, a ring of 2, 3 and 4 slots, frames of two and three macroblock rows, two to
    eight frame slots, and both register budgets of the kernel."""
    n = 9
    frames = [synth.frame(w, h, t, 7, "moving") for t in range(n)]
    types = [0, 1, 1, 1, 0, 1, 1, 1, 1]
    qualities = [16, 16, 8, 8, 24, 24, 16, 31, 1]
    want, wrec = _bins_sequence({"EVXGPU_FRAME_OVERLAP": "0"}, w, h, R, frames, types, qualities)
    for env in ({"EVXGPU_FRAME_OVERLAP": "1", "EVXGPU_FRAME_SLOTS": "2"}, {"EVXGPU_FRAME_OVERLAP": "1", "EVXGPU_FRAME_SLOTS": "3", "EVXGPU_K3_REGS": "1"},
                {"EVXGPU_FRAME_OVERLAP": "1"}, {"EVXGPU_FRAME_OVERLAP": "1", "EVXGPU_FRAME_SLOTS": "8", "EVXGPU_K2_CTAS": "0"}):
        got, grec = _bins_sequence(env, w, h, R, frames, types, qualities)
        for t in range(n):
            assert _same_bins(got[t], want[t]), (env, t)
        assert all((a == b).all() for a, b in zip(grec, wrec)), env
