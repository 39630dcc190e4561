"""ctypes binding of the device C-ABI (include/evxgpu.h).  Fails loudly when the CUDA
library is missing or no device is usable -- there is no CPU fallback in the product."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_SO = os.path.join(HERE, "libevxgpu.so")

BLOCK_DESC_DTYPE = np.dtype({
    "names": ["block_type", "prediction_target", "motion_x", "motion_y", "sp_pred", "sp_amount", "sp_index",
              "q_index", "variance"],
    "formats": ["<i4", "u1", "<i2", "<i2", "u1", "u1", "u1", "u1", "<i2"],
    "offsets": [0, 4, 6, 8, 10, 11, 12, 13, 14],
    "itemsize": 16,
})

T_NAMES = ["convert_in", "inter_search", "wavefront", "deblock", "decode_recon", "convert_out", "bins"]


class Config(C.Structure):
    _fields_ = [("ref_count", C.c_int32), ("linear_quant", C.c_int32), ("deblocking", C.c_int32), ("frame_slots", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(GPU_SO):
        raise RuntimeError(f"{GPU_SO} is missing: run `python -m cairo_b200.build` (no CPU fallback exists)")
    L = C.CDLL(GPU_SO)
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
    L.evxgpu_last_error.restype = C.c_char_p
    L.evxgpu_create.argtypes = [i32, i32, i32, C.POINTER(Config), vp, C.POINTER(vp)]
    L.evxgpu_destroy.argtypes = [vp]
    L.evxgpu_reset.argtypes = [vp]
    L.evxgpu_block_count.argtypes = [vp]
    L.evxgpu_synchronize.argtypes = [vp]
    L.evxgpu_host_alloc.restype = vp
    L.evxgpu_host_alloc.argtypes = [C.c_uint64]
    L.evxgpu_host_free.argtypes = [vp]
    L.evxgpu_device_alloc.restype = vp
    L.evxgpu_device_alloc.argtypes = [C.c_uint64]
    L.evxgpu_device_free.argtypes = [vp]
    L.evxgpu_upload.argtypes = [vp, vp, vp, C.c_uint64]
    L.evxgpu_encode_submit.argtypes = [vp, vp, i32, i32, u32, i32]
    L.evxgpu_encode_collect.argtypes = [vp, vp, vp, C.POINTER(u32)]
    L.evxgpu_encode_capacity.argtypes = [vp]
    L.evxgpu_set_output.argtypes = [vp, i32]
    L.evxgpu_encode_collect_bins.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64), C.POINTER(u32)]
    L.evxgpu_debug_set_bins_capacity.argtypes = [vp, u32]
    L.evxgpu_decode_submit.argtypes = [vp, vp, vp, u32, i32, u32]
    L.evxgpu_decode_collect.argtypes = [vp, vp, i32]
    L.evxgpu_decode_collect_begin.argtypes = [vp, vp, i32]
    L.evxgpu_decode_collect_end.argtypes = [vp]
    L.evxgpu_stage_convert_in.argtypes = [vp, vp]
    L.evxgpu_stage_inter_search.argtypes = [vp, u32, i32]
    L.evxgpu_stage_get_inter_result.argtypes = [vp, i32, vp, vp]
    L.evxgpu_stage_deblock.argtypes = [vp, u32]
    L.evxgpu_stage_set_block_table.argtypes = [vp, vp]
    L.evxgpu_peek_plane.argtypes = [vp, i32, i32, i32, vp]
    L.evxgpu_poke_plane.argtypes = [vp, i32, i32, i32, vp]
    L.evxgpu_get_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.evxgpu_enable_timing.argtypes = [vp, i32]
    L.evxgpu_get_timing_sum.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.evxgpu_get_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), i32]
    L.evxgpu_get_counters_split.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    L.evxgpu_set_encode_grid.argtypes = [vp, i32]
    L.evxgpu_d2h_bytes.restype = C.c_uint64
    L.evxgpu_d2h_bytes.argtypes = [vp]
    L.evxgpu_launch_count.restype = C.c_uint64
    L.evxgpu_launch_count.argtypes = [vp]
    L.evxgpu_time_kernel.restype = C.c_double
    L.evxgpu_time_kernel.argtypes = [vp, i32, i32]
    L.evxgpu_timeline_mark.argtypes = [vp]
    L.evxgpu_last_done_ms.restype = C.c_double
    L.evxgpu_last_done_ms.argtypes = [vp]
    L.evxgpu_measure_int_peak.restype = C.c_double
    L.evxgpu_measure_int_peak.argtypes = [i32, i32]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with status {rc}: {lib().evxgpu_last_error().decode()}")


class Pipeline:
    """One video stream's device state (evxgpu_handle)."""

    def __init__(self, width, height, ref_count=4, linear_quant=0, deblocking=1, device=0, stream=None, frame_slots=0):
        self.L = lib()
        self.w, self.h_ = width, height
        self.aw, self.ah = (width + 15) // 16 * 16, (height + 15) // 16 * 16
        self.R = ref_count
        cfg = Config(ref_count, linear_quant, deblocking, frame_slots)
        h = C.c_void_p()
        _check(self.L.evxgpu_create(device, width, height, C.byref(cfg), stream, C.byref(h)), "evxgpu_create")
        self.h = h
        self.nblocks = self.L.evxgpu_block_count(self.h)
        self._records = np.zeros((self.nblocks, 384), dtype=np.int16)

    def close(self):
        if getattr(self, "h", None):
            self.L.evxgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        _check(self.L.evxgpu_reset(self.h), "evxgpu_reset")

    def encode_submit(self, rgb, frame_type, index, quality):
        """rgb: numpy uint8 (h,w,3) host array, or an int device pointer."""
        if isinstance(rgb, int):
            _check(self.L.evxgpu_encode_submit(self.h, rgb, 1, frame_type, index, quality), "evxgpu_encode_submit")
        else:
            rgb = np.ascontiguousarray(rgb)
            self._keep = rgb
            _check(self.L.evxgpu_encode_submit(self.h, _p(rgb), 0, frame_type, index, quality), "evxgpu_encode_submit")

    def encode_capacity(self):
        """Submitted, uncollected frames the handle accepts (1, or its number of frame slots with bin-string output)."""
        return int(self.L.evxgpu_encode_capacity(self.h))

    def encode_collect(self):
        tbl = np.zeros(self.nblocks, dtype=BLOCK_DESC_DTYPE)
        n = C.c_uint32(0)
        _check(self.L.evxgpu_encode_collect(self.h, _p(tbl), _p(self._records), C.byref(n)), "evxgpu_encode_collect")
        return tbl, self._records[:n.value].copy()

    def set_output(self, mode):
        """0 = table + records, 1 = the slice's bin string only, 2 = both (collect bins first)."""
        _check(self.L.evxgpu_set_output(self.h, mode), "evxgpu_set_output")

    def encode_collect_bins(self):
        """-> (uint64 words of the bin string, number of bins, non-copy macroblocks)"""
        ptr = C.POINTER(C.c_uint64)()
        n, nc = C.c_uint64(0), C.c_uint32(0)
        _check(self.L.evxgpu_encode_collect_bins(self.h, C.byref(ptr), C.byref(n), C.byref(nc)), "evxgpu_encode_collect_bins")
        words = (n.value + 63) // 64
        arr = np.ctypeslib.as_array(ptr, shape=(max(words, 1),)).copy() if words else np.zeros(1, np.uint64)
        return arr, n.value, nc.value

    def set_bins_capacity(self, bits):
        _check(self.L.evxgpu_debug_set_bins_capacity(self.h, bits), "evxgpu_debug_set_bins_capacity")

    def encode(self, rgb, frame_type, index, quality):
        self.encode_submit(rgb, frame_type, index, quality)
        return self.encode_collect()

    def decode(self, table, records, frame_type, index):
        table = np.ascontiguousarray(table)
        records = np.ascontiguousarray(records, dtype=np.int16)
        _check(self.L.evxgpu_decode_submit(self.h, _p(table), _p(records), records.shape[0] if records.size else 0,
                                           frame_type, index), "evxgpu_decode_submit")
        out = np.zeros((self.h_, self.w, 3), dtype=np.uint8)
        _check(self.L.evxgpu_decode_collect(self.h, _p(out), 0), "evxgpu_decode_collect")
        return out

    def convert_in(self, rgb):
        rgb = np.ascontiguousarray(rgb)
        _check(self.L.evxgpu_stage_convert_in(self.h, _p(rgb)), "evxgpu_stage_convert_in")

    def inter_search(self, index, quality):
        _check(self.L.evxgpu_stage_inter_search(self.h, index, quality), "evxgpu_stage_inter_search")

    def inter_result(self, offset):
        d = np.zeros(self.nblocks, dtype=BLOCK_DESC_DTYPE)
        sad = np.zeros(self.nblocks, dtype=np.int32)
        _check(self.L.evxgpu_stage_get_inter_result(self.h, offset, _p(d), _p(sad)), "evxgpu_stage_get_inter_result")
        return d, sad

    def deblock(self, index):
        _check(self.L.evxgpu_stage_deblock(self.h, index), "evxgpu_stage_deblock")

    def set_block_table(self, tbl):
        tbl = np.ascontiguousarray(tbl)
        _check(self.L.evxgpu_stage_set_block_table(self.h, _p(tbl)), "evxgpu_stage_set_block_table")

    def plane(self, which, slot, comp):
        w, h = (self.aw, self.ah) if comp == 0 else (self.aw // 2, self.ah // 2)
        out = np.zeros((h, w), dtype=np.int16)
        _check(self.L.evxgpu_peek_plane(self.h, which, slot, comp, _p(out)), "evxgpu_peek_plane")
        return out

    def planes(self, which, slot=0):
        return [self.plane(which, slot, c) for c in range(3)]

    def set_plane(self, which, slot, comp, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int16)
        _check(self.L.evxgpu_poke_plane(self.h, which, slot, comp, _p(arr)), "evxgpu_poke_plane")

    def enable_timing(self, on=True):
        self.L.evxgpu_enable_timing(self.h, int(on))

    def timing(self):
        ms = (C.c_float * len(T_NAMES))()
        _check(self.L.evxgpu_get_timing(self.h, ms), "evxgpu_get_timing")
        return dict(zip(T_NAMES, [float(v) for v in ms]))

    def timing_sum(self, reset=False):
        """Kernel times (ms) summed over the frames encoded since the last reset."""
        buf = (C.c_double * len(T_NAMES))()
        _check(self.L.evxgpu_get_timing_sum(self.h, buf, 1 if reset else 0), "evxgpu_get_timing_sum")
        return {n: buf[i] for i, n in enumerate(T_NAMES)}

    def counters(self, reset=False):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self.L.evxgpu_get_counters(self.h, C.byref(a), C.byref(b), int(reset)), "evxgpu_get_counters")
        return a.value, b.value

    def counters_split(self, reset=False):
        """(inter full-pel, inter sub-pel, intra full-pel, intra sub-pel) evaluated since the last reset."""
        out = (C.c_uint64 * 4)()
        _check(self.L.evxgpu_get_counters_split(self.h, out, int(reset)), "evxgpu_get_counters_split")
        return tuple(int(v) for v in out)

    def set_encode_grid(self, ctas):
        _check(self.L.evxgpu_set_encode_grid(self.h, int(ctas)), "evxgpu_set_encode_grid")

    def d2h_bytes(self):
        return int(self.L.evxgpu_d2h_bytes(self.h))

    def timeline_mark(self):
        _check(self.L.evxgpu_timeline_mark(self.h), "evxgpu_timeline_mark")

    def last_done_ms(self):
        """Device time (ms since timeline_mark) at which the results of the frame collected last had left the device."""
        return float(self.L.evxgpu_last_done_ms(self.h))

    def launch_count(self):
        return int(self.L.evxgpu_launch_count(self.h))

    def time_kernel(self, kind, reps=50):
        """Average ms of `reps` back-to-back launches of K1 (0), K4 (1) or K6 (2) on this idle handle."""
        return float(self.L.evxgpu_time_kernel(self.h, int(kind), int(reps)))


def records_to_planes(table, records, coef_planes, aw, ah):
    """Scatter the non-copy macroblocks' records into persistent coefficient planes
    (the host mirror of output_cache; copy blocks keep their stale contents)."""
    y, u, v = coef_planes
    mbw = aw // 16
    k = 0
    for mb in range(table.shape[0]):
        if table["block_type"][mb] & 4:
            continue
        bx, by = mb % mbw, mb // mbw
        r = records[k]
        y[by * 16:by * 16 + 16, bx * 16:bx * 16 + 16] = r[:256].reshape(16, 16)
        u[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = r[256:320].reshape(8, 8)
        v[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = r[320:384].reshape(8, 8)
        k += 1
    assert k == records.shape[0]


def planes_to_records(table, coef_planes, aw):
    y, u, v = coef_planes
    mbw = aw // 16
    out = []
    for mb in range(table.shape[0]):
        if table["block_type"][mb] & 4:
            continue
        bx, by = mb % mbw, mb // mbw
        out.append(np.concatenate([y[by * 16:by * 16 + 16, bx * 16:bx * 16 + 16].ravel(),
                                   u[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].ravel(),
                                   v[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8].ravel()]))
    return np.array(out, dtype=np.int16).reshape(-1, 384)
