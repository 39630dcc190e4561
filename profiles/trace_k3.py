"""Row-to-row hand-off latency of evx_wavefront from per-macroblock globaltimer stamps.
Needs a trace build:  EVX_EXTRA_NVCC=-DEVX_K3_TRACE python -m cairo_b200.build --force
stamps per macroblock: [0] start (window up to column n+1 staged), [1] far-right column confirmed
(full2 wait over), [2] end (reconstruction stored).  Not for benchmarking."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H = 1920, 1080
frames = [synth.frame(W, H, t, 0, 'moving') for t in range(4)]
p = gpu.Pipeline(W, H, 2, 0, 1)
for t in range(3):
    p.encode(frames[t], 0 if t == 0 else 1, t, 16)
L.evxgpu_debug_profile(p.h, 1, None)
tbl, rec = p.encode(frames[3], 1, 3, 16)
mbh, mbw = p.ah // 16, p.aw // 16
raw = np.zeros((mbh * 10 + p.nblocks * 4) * 8, dtype=np.int64)      # (eight pipeline frames deep; the stand-alone launch uses the first)
L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
tr = raw[mbh * 10:].reshape(mbh, mbw, 4).astype(np.float64)
if tr[:, :, 0].max() == 0:
    sys.exit("no stamps: build with -DEVX_K3_TRACE")
t0 = tr[0, 0, 0]
start, need2, end = (tr[:, :, 0] - t0) / 1e3, (tr[:, :, 1] - t0) / 1e3, (tr[:, :, 2] - t0) / 1e3
print("globaltimer granularity (ns):", np.unique(np.diff(np.unique(tr[:, :, 0])))[:5])
print("frame: %.1f us" % end.max())
dur = end - start
print("macroblock duration us: mean %.2f median %.2f p90 %.2f" % (dur.mean(), np.median(dur), np.percentile(dur, 90)))
print("row start times (us) rows 0..5:", start[:6, 0].round(1).tolist())
print("row-to-row lag at column 60 (us):", np.diff(start[:, 60])[:12].round(1).tolist(), "mean", np.diff(start[:, 60]).mean().round(2))
# hand-off: producer end of (n+2, by-1) -> consumer's full2 confirmation of (n, by)
h = need2[1:, :mbw - 2] - end[:-1, 2:]
print("hand-off end(n+2,by-1) -> need2(n,by) us: median %.2f p10 %.2f p90 %.2f min %.2f" % (np.median(h), np.percentile(h, 10), np.percentile(h, 90), h.min()))
w = need2 - start
print("time from macroblock start to full2 confirmed us: median %.2f  (macroblock median %.2f)" % (np.median(w), np.median(dur)))
s1 = start[1:, :mbw - 2] - end[:-1, 1:mbw - 1]
print("start(n,by) - end(n+1,by-1) us: median %.2f p10 %.2f" % (np.median(s1), np.percentile(s1, 10)))
gap = start[:, 1:] - end[:, :-1]
print("gap between consecutive macroblocks of a row us: median %.2f p90 %.2f" % (np.median(gap), np.percentile(gap, 90)))
if '-v' in sys.argv:
    np.set_printoptions(linewidth=200, suppress=True)
    for r in range(4):
        print("row", r, "start ", start[r, :8].round(1).tolist())
        print("row", r, "need2 ", need2[r, :8].round(1).tolist())
        print("row", r, "end   ", end[r, :8].round(1).tolist())
    for r in (30, 31):
        print("row", r, "start ", start[r, 50:58].round(1).tolist())
        print("row", r, "need2 ", need2[r, 50:58].round(1).tolist())
        print("row", r, "end   ", end[r, 50:58].round(1).tolist())
    print("types row 31 cols 50..57:", [int(x) for x in tbl['block_type'][31 * mbw + 50: 31 * mbw + 58]])
