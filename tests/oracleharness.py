"""ctypes wrapper for oracle/libevx_oracle.so (plain-C restatement).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libevx_oracle.so")

BLOCK_DESC_DTYPE = np.dtype({
    "names": ["block_type", "prediction_target", "motion_x", "motion_y", "sp_pred", "sp_amount", "sp_index",
              "q_index", "variance"],
    "formats": ["<i4", "u1", "<i2", "<i2", "u1", "u1", "u1", "u1", "<i2"],
    "offsets": [0, 4, 6, 8, 10, 11, 12, 13, 14],
    "itemsize": 16,
})

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("evx_oracle.c", "evx_oracle.h"))
    if not os.path.exists(SO) or os.path.getmtime(SO) < src_m:
        build()
    L = C.CDLL(SO)
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
    L.evxo_create.restype = vp
    L.evxo_create.argtypes = [i32] * 5
    L.evxo_destroy.argtypes = [vp]
    L.evxo_reset.argtypes = [vp]
    L.evxo_block_count.argtypes = [vp]
    L.evxo_plane.restype = C.POINTER(C.c_int16)
    L.evxo_plane.argtypes = [vp, i32, i32, i32]
    L.evxo_block_table.restype = vp
    L.evxo_block_table.argtypes = [vp]
    L.evxo_convert_in.argtypes = [vp, vp]
    L.evxo_encode_slice.argtypes = [vp, i32, u32, i32]
    L.evxo_encode_slice_wavefront.argtypes = [vp, i32, u32, i32, i32]
    L.evxo_decode_slice.argtypes = [vp, i32, u32]
    L.evxo_deblock.argtypes = [vp, u32]
    L.evxo_deblock_tiled.argtypes = [vp, u32]
    L.evxo_convert_out.argtypes = [vp, u32, vp]
    L.evxo_serialize_slice.restype = u32
    L.evxo_serialize_slice.argtypes = [vp, vp, u32]
    L.evxo_unserialize_slice.argtypes = [vp, vp, u32]
    L.evxo_inter_prediction.restype = C.c_int32
    L.evxo_inter_prediction.argtypes = [vp, u32, i32, i32, i32, i32, vp]
    L.evxo_intra_prediction.restype = C.c_int32
    L.evxo_intra_prediction.argtypes = [vp, u32, i32, i32, i32, vp]
    L.evxo_get_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.evxo_reset_counters.argtypes = [vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, width, height, ref_count=4, linear_quant=0, deblocking=1):
        self.L = lib()
        self.w, self.h_ = width, height
        self.aw, self.ah = (width + 15) // 16 * 16, (height + 15) // 16 * 16
        self.R = ref_count
        self.h = self.L.evxo_create(width, height, ref_count, linear_quant, deblocking)
        assert self.h
        self.nblocks = self.L.evxo_block_count(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.evxo_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.evxo_reset(self.h)

    def plane(self, which, slot, comp):
        """numpy VIEW onto the oracle's plane (writes go through)."""
        w, h = (self.aw, self.ah) if comp == 0 else (self.aw // 2, self.ah // 2)
        ptr = self.L.evxo_plane(self.h, which, slot, comp)
        return np.ctypeslib.as_array(ptr, shape=(h, w))

    def planes(self, which, slot=0):
        return [self.plane(which, slot, c).copy() for c in range(3)]

    def block_table(self):
        ptr = self.L.evxo_block_table(self.h)
        buf = (C.c_uint8 * (16 * self.nblocks)).from_address(ptr)
        return np.frombuffer(buf, dtype=BLOCK_DESC_DTYPE)

    def convert_in(self, rgb):
        rgb = np.ascontiguousarray(rgb)
        self.L.evxo_convert_in(self.h, _p(rgb))

    def encode_slice(self, ftype, index, quality, wavefront=None):
        if wavefront is None:
            self.L.evxo_encode_slice(self.h, ftype, index, quality)
        else:
            self.L.evxo_encode_slice_wavefront(self.h, ftype, index, quality, int(wavefront))

    def decode_slice(self, ftype, index):
        self.L.evxo_decode_slice(self.h, ftype, index)

    def deblock(self, index, tiled=False):
        (self.L.evxo_deblock_tiled if tiled else self.L.evxo_deblock)(self.h, index)

    def convert_out(self, index):
        out = np.zeros((self.h_, self.w, 3), dtype=np.uint8)
        self.L.evxo_convert_out(self.h, index, _p(out))
        return out

    def serialize(self):
        cap = self.aw * self.ah * 6 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        bits = self.L.evxo_serialize_slice(self.h, _p(out), cap)
        return out[:(bits + 7) // 8].copy(), bits

    def unserialize(self, data, nbits):
        data = np.ascontiguousarray(data)
        self.L.evxo_unserialize_slice(self.h, _p(data), nbits)

    def inter_prediction(self, index, quality, px, py, offset):
        d = np.zeros(1, dtype=BLOCK_DESC_DTYPE)
        sad = self.L.evxo_inter_prediction(self.h, index, quality, px, py, offset, _p(d))
        return d[0], sad

    def intra_prediction(self, index, quality, px, py):
        d = np.zeros(1, dtype=BLOCK_DESC_DTYPE)
        sad = self.L.evxo_intra_prediction(self.h, index, quality, px, py, _p(d))
        return d[0], sad

    def counters(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.L.evxo_get_counters(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def bits_equal(a, abits, b, bbits):
    """Compare two LSB-first bit strings at bit length."""
    if abits != bbits:
        return False
    ua = np.unpackbits(np.asarray(a, dtype=np.uint8), bitorder="little")[:abits]
    ub = np.unpackbits(np.asarray(b, dtype=np.uint8), bitorder="little")[:bbits]
    return bool((ua == ub).all())


def tables_equal(a, b, check_variance=True):
    """Field-wise block-table comparison; q_index/variance only where defined (non-copy blocks),
    motion/sub-pel fields only where the type uses them."""
    if not (a["block_type"] == b["block_type"]).all():
        return False
    t = a["block_type"]
    intra, motion, copy = (t & 1) != 0, (t & 2) != 0, (t & 4) != 0
    ok = (a["prediction_target"][~intra] == b["prediction_target"][~intra]).all()
    for f in ("motion_x", "motion_y", "sp_pred"):
        ok &= (a[f][motion] == b[f][motion]).all()
    sp = motion & (a["sp_pred"] != 0)
    for f in ("sp_amount", "sp_index"):
        ok &= (a[f][sp] == b[f][sp]).all()
    for f in ("q_index", "variance") if check_variance else ("q_index",):
        ok &= (a[f][~copy] == b[f][~copy]).all()
    return bool(ok)
