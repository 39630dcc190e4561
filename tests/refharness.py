"""ctypes wrapper for oracle/_ref/libevxref_<variant>.so (the unmodified
reference, compiled by oracle/Makefile).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

BLOCK_DESC_DTYPE = np.dtype({
    "names": ["block_type", "prediction_target", "motion_x", "motion_y", "sp_pred", "sp_amount", "sp_index",
              "q_index", "variance"],
    "formats": ["<i4", "u1", "<i2", "<i2", "u1", "u1", "u1", "u1", "<i2"],
    "offsets": [0, 4, 6, 8, 10, 11, 12, 13, 14],
    "itemsize": 16,
})


def available(variant="r4"):
    return os.path.exists(os.path.join(REF_DIR, f"libevxref_{variant}.so"))


_libs = {}


def lib(variant="r4"):
    if variant in _libs:
        return _libs[variant]
    L = C.CDLL(os.path.join(REF_DIR, f"libevxref_{variant}.so"))
    vp, u8p, u32, i32 = C.c_void_p, C.c_void_p, C.c_uint32, C.c_int
    L.evxref_encoder_create.restype = vp
    L.evxref_encoder_destroy.argtypes = [vp]
    L.evxref_encoder_clear.argtypes = [vp]
    L.evxref_encoder_insert_intra.argtypes = [vp]
    L.evxref_encoder_set_quality.argtypes = [vp, i32]
    L.evxref_encoder_encode.argtypes = [vp, u8p, u32, u32, u8p, u32, C.POINTER(u32)]
    if hasattr(L, "evxref_encoder_peek"):
        L.evxref_encoder_peek.argtypes = [vp, i32, u8p]
    L.evxref_decoder_create.restype = vp
    L.evxref_decoder_destroy.argtypes = [vp]
    L.evxref_decoder_clear.argtypes = [vp]
    L.evxref_decoder_decode.argtypes = [vp, u8p, u32, u8p]
    L.evxref_stage_create.restype = vp
    L.evxref_stage_create.argtypes = [u32, u32]
    L.evxref_stage_destroy.argtypes = [vp]
    L.evxref_stage_set_frame.argtypes = [vp, i32, u32, i32]
    L.evxref_stage_convert_in.argtypes = [vp, u8p]
    L.evxref_stage_encode_slice.argtypes = [vp]
    L.evxref_stage_decode_slice.argtypes = [vp]
    L.evxref_stage_deblock.argtypes = [vp]
    L.evxref_stage_serialize.argtypes = [vp, u8p, u32, C.POINTER(u32)]
    L.evxref_stage_unserialize.argtypes = [vp, u8p, u32]
    L.evxref_stage_convert_out.argtypes = [vp, u8p]
    L.evxref_stage_inter_prediction.argtypes = [vp, i32, i32, i32, vp, C.POINTER(C.c_int32)]
    L.evxref_stage_intra_prediction.argtypes = [vp, i32, i32, vp, C.POINTER(C.c_int32)]
    L.evxref_stage_get_plane.argtypes = [vp, i32, i32, i32, vp]
    L.evxref_stage_set_plane.argtypes = [vp, i32, i32, i32, vp]
    L.evxref_stage_get_block_table.argtypes = [vp, vp]
    L.evxref_stage_set_block_table.argtypes = [vp, vp]
    _libs[variant] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefEncoder:
    """evx1_encoder through the reference's public API (evx1.h:66-94)."""

    def __init__(self, variant="r4"):
        self.L = lib(variant)
        self.h = self.L.evxref_encoder_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evxref_encoder_destroy(self.h)
            self.h = None

    def set_quality(self, q):
        return self.L.evxref_encoder_set_quality(self.h, q)

    def insert_intra(self):
        return self.L.evxref_encoder_insert_intra(self.h)

    def clear(self):
        return self.L.evxref_encoder_clear(self.h)

    def encode(self, rgb):
        h, w, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb)
        cap = w * h * 6 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_uint32(0)
        st = self.L.evxref_encoder_encode(self.h, _p(rgb), w, h, _p(out), cap, C.byref(bits))
        assert st == 0, st
        return out[:(bits.value + 7) // 8].copy(), bits.value


    def peek(self, state, width, height):
        out = np.zeros((height, width, 3), dtype=np.uint8)
        st = self.L.evxref_encoder_peek(self.h, int(state), _p(out))
        assert st == 0, st
        return out


class RefDecoder:
    def __init__(self, variant="r4"):
        self.L = lib(variant)
        self.h = self.L.evxref_decoder_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evxref_decoder_destroy(self.h)
            self.h = None

    def decode(self, data, nbits, width, height):
        out = np.zeros((height, width, 3), dtype=np.uint8)
        data = np.ascontiguousarray(data)
        st = self.L.evxref_decoder_decode(self.h, _p(data), nbits, _p(out))
        assert st == 0, st
        return out


class RefStage:
    """Stage-by-stage access to the reference's engine (encode.cpp:205-232, decode.cpp:172-198)."""

    def __init__(self, width, height, variant="r4"):
        self.L = lib(variant)
        self.w, self.h_ = width, height
        self.aw, self.ah = (width + 15) // 16 * 16, (height + 15) // 16 * 16
        self.R = self.L.evxref_ref_count()
        self.h = self.L.evxref_stage_create(width, height)
        self.nblocks = (self.aw // 16) * (self.ah // 16)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evxref_stage_destroy(self.h)
            self.h = None

    def set_frame(self, ftype, index, quality):
        self.L.evxref_stage_set_frame(self.h, ftype, index, quality)

    def convert_in(self, rgb):
        rgb = np.ascontiguousarray(rgb)
        assert self.L.evxref_stage_convert_in(self.h, _p(rgb)) == 0

    def encode_slice(self):
        assert self.L.evxref_stage_encode_slice(self.h) == 0

    def decode_slice(self):
        assert self.L.evxref_stage_decode_slice(self.h) == 0

    def deblock(self):
        assert self.L.evxref_stage_deblock(self.h) == 0

    def serialize(self):
        cap = self.aw * self.ah * 6 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_uint32(0)
        assert self.L.evxref_stage_serialize(self.h, _p(out), cap, C.byref(bits)) == 0
        return out[:(bits.value + 7) // 8].copy(), bits.value

    def unserialize(self, data, nbits):
        data = np.ascontiguousarray(data)
        assert self.L.evxref_stage_unserialize(self.h, _p(data), nbits) == 0

    def convert_out(self):
        out = np.zeros((self.h_, self.w, 3), dtype=np.uint8)
        assert self.L.evxref_stage_convert_out(self.h, _p(out)) == 0
        return out

    def plane(self, which, slot, comp):
        w, h = (self.aw, self.ah) if comp == 0 else (self.aw // 2, self.ah // 2)
        out = np.zeros((h, w), dtype=np.int16)
        assert self.L.evxref_stage_get_plane(self.h, which, slot, comp, _p(out)) == 0
        return out

    def planes(self, which, slot=0):
        return [self.plane(which, slot, c) for c in range(3)]

    def set_plane(self, which, slot, comp, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int16)
        assert self.L.evxref_stage_set_plane(self.h, which, slot, comp, _p(arr)) == 0

    def block_table(self):
        out = np.zeros(self.nblocks, dtype=BLOCK_DESC_DTYPE)
        self.L.evxref_stage_get_block_table(self.h, _p(out))
        return out

    def set_block_table(self, tbl):
        tbl = np.ascontiguousarray(tbl)
        self.L.evxref_stage_set_block_table(self.h, _p(tbl))

    def inter_prediction(self, px, py, offset):
        d = np.zeros(1, dtype=BLOCK_DESC_DTYPE)
        sad = C.c_int32(0)
        self.L.evxref_stage_inter_prediction(self.h, px, py, offset, _p(d), C.byref(sad))
        return d[0], sad.value

    def intra_prediction(self, px, py):
        d = np.zeros(1, dtype=BLOCK_DESC_DTYPE)
        sad = C.c_int32(0)
        self.L.evxref_stage_intra_prediction(self.h, px, py, _p(d), C.byref(sad))
        return d[0], sad.value
