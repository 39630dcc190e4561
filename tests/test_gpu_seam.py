"""Seam B of INTEGRATION.md run for real: the device library under the REFERENCE's own host entropy stage.

  encoder side   evxgpu_encode_collect (table + records) -> include/evxgpu_records.h scatter -> the reference's
                 context (block_table, cache_bank.output_cache) -> the reference's serialize_slice
                 (serialize.cpp:319-340) -- must give the bytes evx1_encoder::encode of this build gives;
  decoder side   the reference's unserialize_slice (unserialize.cpp:321-341) -> block_table + cache_bank.input_cache
                 -> gather -> evxgpu_decode_submit / collect -- must give the picture the reference's decoder gives.
And the device-resident frame source / sink of the public API."""
import ctypes as C

import numpy as np
import pytest

import refharness as R
from cairo_b200 import synth

pytestmark = pytest.mark.gpu


def _payload(data, bits, first):
    """the slice of one frame: the stream minus the 14-byte header (first frame) and the 10-byte frame descriptor"""
    skip = 24 if first else 10
    return np.ascontiguousarray(data[skip:]), bits - 8 * skip


def _bits_equal(a, abits, b, bbits):
    if abits != bbits:
        return False
    ua = np.unpackbits(np.asarray(a, np.uint8), bitorder="little")[:abits]
    ub = np.unpackbits(np.asarray(b, np.uint8), bitorder="little")[:bbits]
    return bool((ua == ub).all())


@pytest.mark.parametrize("w,h,n", [(352, 288, 8), (1920, 1080, 4)])
def test_device_records_through_the_reference_serialiser(w, h, n):
    if not R.available("r4"):
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    from cairo_b200 import api, gpu
    q, ring = 16, 4
    p = gpu.Pipeline(w, h, ring, 0, 1)                       # table + records output: the seam a reference maintainer binds
    ours = api.evx1_encoder(ref_count=ring)
    ours.set_quality(q)
    rs = R.RefStage(w, h, "r4")
    planes = [np.zeros((rs.ah, rs.aw), np.int16), np.zeros((rs.ah // 2, rs.aw // 2), np.int16), np.zeros((rs.ah // 2, rs.aw // 2), np.int16)]
    for t in range(n):
        f = synth.frame(w, h, t, 4, "moving")
        intra = t == 0 or t == 5
        if intra and t:
            ours.insert_intra()
        tbl, rec = p.encode(f, 0 if intra else 1, t, q)
        assert api.scatter_records(tbl, rec, planes, rs.aw, rs.ah) == rec.shape[0]
        for c in range(3):
            rs.set_plane(1, 0, c, planes[c])                 # cache_bank.output_cache
        rs.set_block_table(tbl)
        rs.set_frame(0 if intra else 1, t, q)
        rd, rb = rs.serialize()                              # the reference's serialize_slice
        data, bits = ours.encode(f)
        pd, pb = _payload(data, bits, t == 0)
        assert _bits_equal(pd, pb, rd, rb), (t, pb, rb)


@pytest.mark.parametrize("w,h,n", [(352, 288, 8), (1920, 1080, 4)])
def test_reference_unserialiser_into_device_decoder(w, h, n):
    if not R.available("r4"):
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    from cairo_b200 import api, gpu
    q, ring = 16, 4
    enc = R.RefEncoder("r4")
    enc.set_quality(q)
    rdec = R.RefDecoder("r4")
    rs = R.RefStage(w, h, "r4")
    p = gpu.Pipeline(w, h, ring, 0, 1)
    for t in range(n):
        f = synth.frame(w, h, t, 6, "moving")
        intra = t == 0 or t == 3
        if intra and t:
            enc.insert_intra()
        data, bits = enc.encode(f)
        want = rdec.decode(data, bits, w, h)
        pd, pb = _payload(data, bits, t == 0)
        rs.set_frame(0 if intra else 1, t, q)
        rs.unserialize(pd, pb)                               # the reference's unserialize_slice
        tbl = rs.block_table()
        rec = api.gather_records(tbl, rs.planes(0), rs.aw, rs.ah)      # cache_bank.input_cache
        rgb = p.decode(tbl, rec, 0 if intra else 1, t)
        assert (rgb == want).all(), t


def test_device_resident_frames_through_the_public_api():
    """evx1_config::device_frames: encode()/submit() read RGB8 frames from device memory, decode()/collect() leave the
    picture in device memory; streams and pictures equal the host-frame run."""
    import torch
    from cairo_b200 import api
    w, h, n, q = 1920, 1080, 8, 16
    frames = [synth.frame(w, h, t, 2, "moving") for t in range(n)]
    a = api.evx1_encoder(ref_count=2)
    a.set_quality(q)
    dec_host = api.evx1_decoder()
    want, pics = [], []
    for t in range(n):
        d, b = a.encode(frames[t])
        want.append((d.copy(), b))
        pics.append(dec_host.decode(d, b, w, h).copy())
    dev = torch.from_numpy(np.stack(frames)).cuda()
    e = api.evx1_encoder(ref_count=2, device_frames=True)
    e.set_quality(q)
    got = []
    d, b = e.encode((int(dev[0].data_ptr()), w, h))
    got.append((d.copy(), b))
    look = 4
    for t in range(1, n):
        e.submit((int(dev[t].data_ptr()), w, h))
        if t > look:
            d, b = e.collect()
            got.append((d.copy(), b))
    while len(got) < n:
        d, b = e.collect()
        got.append((d.copy(), b))
    for t in range(n):
        assert got[t][1] == want[t][1] and (got[t][0] == want[t][0]).all(), t
    dec = api.evx1_decoder(device_frames=True)
    out = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    for t in range(n):
        dec.decode(want[t][0], want[t][1], w, h, out=int(out.data_ptr()))
        torch.cuda.synchronize()
        assert (out.cpu().numpy() == pics[t]).all(), t


def test_decode_two_frames_on_the_device():
    """include/evxgpu.h, evxgpu_decode_collect_begin / _end: once the copy-out of frame n has begun, frame n+1 may be
    submitted (its staging copies and kernels run under that copy).  Same pictures as one frame at a time; a second
    submit before the copy-out has begun, and a third frame, are refused with status 8."""
    import ctypes as C
    from cairo_b200 import gpu
    w, h, q, n = 352, 288, 16, 7
    frames = [synth.frame(w, h, t, 3, "moving") for t in range(n)]
    enc = gpu.Pipeline(w, h, 2, 0, 1)
    coded = []
    for t in range(n):
        tbl, rec = enc.encode(frames[t], 0 if t == 0 else 1, t, q)
        coded.append((tbl.copy(), rec.copy()))
    one = gpu.Pipeline(w, h, 2, 0, 1)
    want = [one.decode(tbl, rec, 0 if t == 0 else 1, t).copy() for t, (tbl, rec) in enumerate(coded)]

    L = gpu.lib()
    two = gpu.Pipeline(w, h, 2, 0, 1)
    outs = [np.zeros((h, w, 3), dtype=np.uint8) for _ in range(n)]

    def submit(t):
        tbl, rec = coded[t]
        tbl = np.ascontiguousarray(tbl); rec = np.ascontiguousarray(rec, dtype=np.int16)
        return L.evxgpu_decode_submit(two.h, tbl.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p), rec.shape[0] if rec.size else 0,
                                      0 if t == 0 else 1, t)

    assert L.evxgpu_decode_collect_end(two.h) == 15                       # nothing submitted
    assert submit(0) == 0
    assert submit(1) == 8                                                  # frame 0's copy-out has not begun
    for t in range(n):
        assert L.evxgpu_decode_collect_begin(two.h, outs[t].ctypes.data_as(C.c_void_p), 0) == 0
        assert L.evxgpu_decode_collect_begin(two.h, outs[t].ctypes.data_as(C.c_void_p), 0) == 8      # begun already
        if t + 1 < n:
            assert submit(t + 1) == 0                                      # under frame t's copy-out
            assert submit(t + 1) == 8                                      # a third frame
        assert L.evxgpu_decode_collect_end(two.h) == 0
    assert L.evxgpu_decode_collect_begin(two.h, outs[0].ctypes.data_as(C.c_void_p), 0) == 15
    for t in range(n):
        assert (outs[t] == want[t]).all(), t
