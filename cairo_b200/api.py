"""Python mirror of the reference's public interface (evx1.h:66-122) over the C++ host library
(cairo_b200/csrc/host, bound through include/evx1_c.h).  Same names and argument meaning as
evx1_encoder / evx1_decoder: encode() takes an RGB8 frame and returns the bits it appended,
decode() takes one frame's bits and returns RGB8."""
import ctypes as C
import os

import numpy as np

from . import gpu as _gpu

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(HERE, "libevx1.so")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_SO) or not os.path.exists(_gpu.GPU_SO):
        raise RuntimeError("native libraries missing: run `python -m cairo_b200.build` (there is no CPU fallback)")
    C.CDLL(_gpu.GPU_SO, mode=C.RTLD_GLOBAL)
    L = C.CDLL(HOST_SO)
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
    L.evx1c_encoder_create.restype = vp
    L.evx1c_encoder_create.argtypes = [i32] * 6
    L.evx1c_encoder_create_ex.restype = vp
    L.evx1c_encoder_create_ex.argtypes = [i32] * 9
    L.evx1c_encoder_destroy.argtypes = [vp]
    L.evx1c_encoder_clear.argtypes = [vp]
    L.evx1c_encoder_insert_intra.argtypes = [vp]
    L.evx1c_encoder_set_quality.argtypes = [vp, i32]
    L.evx1c_encoder_encode.argtypes = [vp, vp, u32, u32, vp, u32, C.POINTER(u32)]
    L.evx1c_encoder_submit.argtypes = [vp, vp, u32, u32]
    L.evx1c_encoder_collect.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.evx1c_encoder_peek.argtypes = [vp, i32, vp]
    L.evx1c_encoder_wait_ms.restype = C.c_double
    L.evx1c_encoder_wait_ms.argtypes = [vp]
    L.evx1c_encoder_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.evx1c_decoder_create.restype = vp
    L.evx1c_decoder_create.argtypes = [i32] * 3
    L.evx1c_decoder_create_ex.restype = vp
    L.evx1c_decoder_create_ex.argtypes = [i32] * 4
    L.evx1c_decoder_destroy.argtypes = [vp]
    L.evx1c_decoder_clear.argtypes = [vp]
    L.evx1c_decoder_decode.argtypes = [vp, vp, u32, vp]
    L.evx1c_decoder_submit.argtypes = [vp, vp, u32]
    L.evx1c_decoder_collect.argtypes = [vp, vp]
    L.evx1c_decoder_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.evx1c_slice_writer_create.restype = vp
    L.evx1c_slice_writer_create.argtypes = [i32] * 3
    L.evx1c_slice_writer_destroy.argtypes = [vp]
    L.evx1c_slice_writer_serialize.argtypes = [vp, vp, vp, u32, vp, u32, C.POINTER(u32)]
    L.evx1c_slice_writer_serialize_bins.argtypes = [vp, vp, C.c_uint64, vp, u32, C.POINTER(u32)]
    L.evx1c_slice_reader_create.restype = vp
    L.evx1c_slice_reader_create.argtypes = [i32] * 3
    L.evx1c_slice_reader_destroy.argtypes = [vp]
    L.evx1c_slice_reader_unserialize.argtypes = [vp, vp, u32, vp, vp, C.POINTER(u32)]
    L.evx1c_parsed_slice_create.restype = vp
    L.evx1c_parsed_slice_destroy.argtypes = [vp]
    L.evx1c_slice_reader_parse.argtypes = [vp, vp, u32, vp]
    L.evx1c_slice_reader_apply.argtypes = [vp, vp, vp, vp, C.POINTER(u32)]
    L.evx1c_scatter_records.restype = u32
    L.evx1c_scatter_records.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.evx1c_gather_records.restype = u32
    L.evx1c_gather_records.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class evx1_encoder:
    """evx1_encoder (evx1.h:66-94).  ref_count/linear_quant/deblocking default to config.h's values."""

    def __init__(self, device=0, ref_count=-1, linear_quant=-1, deblocking=-1, periodic_intra=-1, default_quality=-1,
                 frame_slots=0, coder_threads=0, device_frames=False):
        """frame_slots / coder_threads: evx1_config additions (0 = defaults).  device_frames=True: the frames given to
        encode()/submit() are device pointers, passed as (ptr, width, height)."""
        self.L = lib()
        self.h = self.L.evx1c_encoder_create_ex(device, ref_count, linear_quant, deblocking, periodic_intra, default_quality,
                                                frame_slots, coder_threads, 1 if device_frames else 0)
        if not self.h:
            raise RuntimeError("create_encoder failed")
        self._out = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evx1c_encoder_destroy(self.h)
            self.h = None

    def clear(self):
        return self.L.evx1c_encoder_clear(self.h)

    def insert_intra(self):
        return self.L.evx1c_encoder_insert_intra(self.h)

    def set_quality(self, quality):
        return self.L.evx1c_encoder_set_quality(self.h, int(quality))

    def _frame(self, image):
        if isinstance(image, tuple):
            ptr, w, h = image
        else:
            image = np.ascontiguousarray(image)
            h, w, _ = image.shape
            ptr = _p(image)
            self._keep = (getattr(self, "_keep", (None, None, None))[1:] + (image,))      # submit(): the last three frames outlive their calls
        cap = w * h * 6 + 4096
        if self._out is None or self._out.size != cap:
            self._out = np.zeros(cap, dtype=np.uint8)
        return ptr, w, h, cap

    def encode(self, image, out=None):
        """image: uint8 (height, width, 3) R8G8B8, host memory (numpy array or a raw pointer with
        width/height given through `image=(ptr, width, height)`).  Returns (bytes, nbits)."""
        ptr, w, h, cap = self._frame(image)
        bits = C.c_uint32(0)
        st = self.L.evx1c_encoder_encode(self.h, ptr, w, h, _p(self._out), cap, C.byref(bits))
        if st != 0:
            raise RuntimeError(f"evx1_encoder::encode failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")
        return self._out[:(bits.value + 7) // 8], bits.value

    def submit(self, image):
        """First half of encode(): queue the frame on the device.  See evx1.h for the pairing rules."""
        ptr, w, h, _ = self._frame(image)
        st = self.L.evx1c_encoder_submit(self.h, ptr, w, h)
        if st != 0:
            raise RuntimeError(f"evx1_encoder::submit failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")

    def collect(self):
        """Second half of encode(): (bytes, nbits) of the oldest uncollected frame."""
        if self._out is None:
            raise RuntimeError("evx1_encoder::collect failed with status 15: nothing submitted")
        bits = C.c_uint32(0)
        st = self.L.evx1c_encoder_collect(self.h, _p(self._out), self._out.size, C.byref(bits))
        if st != 0:
            raise RuntimeError(f"evx1_encoder::collect failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")
        return self._out[:(bits.value + 7) // 8], bits.value

    PEEK_SOURCE, PEEK_PREDICTION, PEEK_BLOCK_TABLE, PEEK_QUANT_TABLE, PEEK_SPMP_TABLE, PEEK_BLOCK_VARIANCE, PEEK_DESTINATION = range(7)

    def peek(self, state, width, height):
        """evx1_encoder::peek: an R8G8B8 debug picture (EVX_PEEK_STATE, evx1.h:53-62)."""
        out = np.zeros((height, width, 3), dtype=np.uint8)
        st = self.L.evx1c_encoder_peek(self.h, int(state), _p(out))
        if st != 0:
            raise RuntimeError(f"evx1_encoder::peek failed with status {st}")
        return out

    def stats(self):
        g, e, b, n, d = C.c_double(0), C.c_double(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self.L.evx1c_encoder_stats(self.h, C.byref(g), C.byref(e), C.byref(b), C.byref(n), C.byref(d))
        return {"gpu_ms": g.value, "entropy_ms": e.value, "slice_bits": b.value, "noncopy_blocks": n.value, "d2h_bytes": d.value,
                "wait_ms": self.L.evx1c_encoder_wait_ms(self.h)}


class evx1_decoder:
    """evx1_decoder (evx1.h:96-113)."""

    def __init__(self, device=0, linear_quant=-1, deblocking=-1, device_frames=False):
        """device_frames=True: decode()/collect() write the picture to a device pointer given as out=ptr (int)."""
        self.L = lib()
        self.device_frames = bool(device_frames)
        self.h = self.L.evx1c_decoder_create_ex(device, linear_quant, deblocking, 1 if device_frames else 0)
        if not self.h:
            raise RuntimeError("create_decoder failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evx1c_decoder_destroy(self.h)
            self.h = None

    def clear(self):
        return self.L.evx1c_decoder_clear(self.h)

    def stats(self):
        g, e = C.c_double(0), C.c_double(0)
        self.L.evx1c_decoder_stats(self.h, C.byref(g), C.byref(e))
        return {"gpu_ms": g.value, "entropy_ms": e.value}

    def decode(self, data, nbits, width, height, out=None):
        """out: optional uint8 (height, width, 3) array to decode into (e.g. a view of pinned memory)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if out is None:
            out = np.empty((height, width, 3), dtype=np.uint8)
        st = self.L.evx1c_decoder_decode(self.h, _p(data), nbits, out if isinstance(out, int) else _p(out))
        if st != 0:
            raise RuntimeError(f"evx1_decoder::decode failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")
        return out


    def submit(self, data, nbits):
        """First half of decode(): entropy-decode one frame on the host and queue it for the device."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        st = self.L.evx1c_decoder_submit(self.h, _p(data), nbits)
        if st != 0:
            raise RuntimeError(f"evx1_decoder::submit failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")

    def collect(self, width, height, out=None):
        """Second half of decode(): the oldest submitted frame's picture."""
        if out is None:
            out = np.empty((height, width, 3), dtype=np.uint8)
        st = self.L.evx1c_decoder_collect(self.h, out if isinstance(out, int) else _p(out))
        if st != 0:
            raise RuntimeError(f"evx1_decoder::collect failed with status {st}: {_gpu.lib().evxgpu_last_error().decode()}")
        return out


class SliceWriter:
    def __init__(self, mbw, mbh, ref_count):
        self.L = lib()
        self.n = mbw * mbh
        self.h = self.L.evx1c_slice_writer_create(mbw, mbh, ref_count)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evx1c_slice_writer_destroy(self.h)
            self.h = None

    def serialize(self, table, records):
        table = np.ascontiguousarray(table)
        records = np.ascontiguousarray(records, dtype=np.int16).reshape(-1, 384)
        cap = self.n * 384 * 5 + 4096
        if getattr(self, "_out", None) is None:
            self._out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_uint32(0)
        st = self.L.evx1c_slice_writer_serialize(self.h, _p(table), _p(records), records.shape[0], _p(self._out), cap, C.byref(bits))
        assert st == 0, st
        return self._out[:(bits.value + 7) // 8].copy(), bits.value


    def serialize_bins(self, words, nbins):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        cap = self.n * 384 * 5 + 4096
        if getattr(self, "_out", None) is None:
            self._out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_uint32(0)
        st = self.L.evx1c_slice_writer_serialize_bins(self.h, _p(words), nbins, _p(self._out), cap, C.byref(bits))
        assert st == 0, st
        return self._out[:(bits.value + 7) // 8].copy(), bits.value


class SliceReader:
    def __init__(self, mbw, mbh, ref_count):
        self.L = lib()
        self.n = mbw * mbh
        self.h = self.L.evx1c_slice_reader_create(mbw, mbh, ref_count)
        self.table = np.zeros(self.n, dtype=_gpu.BLOCK_DESC_DTYPE)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evx1c_slice_reader_destroy(self.h)
            self.h = None

    def parse(self, data, nbits):
        """State-free half: returns an opaque parsed slice (free it with free_parsed)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        ps = self.L.evx1c_parsed_slice_create()
        st = self.L.evx1c_slice_reader_parse(self.h, _p(data), nbits, ps)
        assert st == 0, st
        return ps

    def apply(self, ps):
        """In-order half: merges a parsed slice into the stream's state; returns (table, records)."""
        rec = np.zeros((self.n, 384), dtype=np.int16)
        n = C.c_uint32(0)
        st = self.L.evx1c_slice_reader_apply(self.h, ps, _p(self.table), _p(rec), C.byref(n))
        assert st == 0, st
        return self.table.copy(), rec[:n.value].copy()

    def free_parsed(self, ps):
        self.L.evx1c_parsed_slice_destroy(ps)

    def unserialize(self, data, nbits):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        rec = np.zeros((self.n, 384), dtype=np.int16)
        n = C.c_uint32(0)
        st = self.L.evx1c_slice_reader_unserialize(self.h, _p(data), nbits, _p(self.table), _p(rec), C.byref(n))
        assert st == 0, st
        return self.table.copy(), rec[:n.value].copy()


def scatter_records(table, records, planes, aw, ah):
    """include/evxgpu_records.h: the non-copy macroblocks' records into persistent coefficient planes (Y, U, V int16)."""
    table = np.ascontiguousarray(table)
    records = np.ascontiguousarray(records, dtype=np.int16)
    y, u, v = planes
    return int(lib().evx1c_scatter_records(_p(table), _p(records), aw, ah, _p(y), _p(u), _p(v)))


def gather_records(table, planes, aw, ah):
    """include/evxgpu_records.h: the records of the table's non-copy macroblocks out of coefficient planes."""
    table = np.ascontiguousarray(table)
    y, u, v = [np.ascontiguousarray(a, dtype=np.int16) for a in planes]
    out = np.zeros((table.shape[0], 384), dtype=np.int16)
    n = int(lib().evx1c_gather_records(_p(table), _p(y), _p(u), _p(v), aw, ah, _p(out)))
    return out[:n].copy()
