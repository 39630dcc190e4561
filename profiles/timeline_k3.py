"""Stage-by-stage cycle timeline of one full-pel search round of evx_wavefront (debug build with
clock64 probes; see the git history of evx_wavefront.cuh for the probe points)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H = 1920, 16
frames = [np.ascontiguousarray(synth.frame(W, 1080, t, 0, 'moving')[500:516]) for t in range(4)]
p = gpu.Pipeline(W, H, 2, 0, 1)
for t in range(3): p.encode(frames[t], 0 if t == 0 else 1, t, 16)
L.evxgpu_debug_profile(p.h, 1, None)
p.encode(frames[3], 1, 3, 16)
prof = np.zeros(1*10 + 68*10 + 128, dtype=np.int64)
L.evxgpu_debug_profile(p.h, 1, prof.ctypes.data_as(C.c_void_p))
d = prof[680:680+96].reshape(8, 12)
t0 = d[:, 0].min()
print("warp: [start, lanecost_done, redux_done, published, prebarrier_done, first_read, selected, updated]")
for w in range(8):
    r = d[w, :8] - t0
    print(w, r.tolist(), " deltas:", np.diff(r).tolist())
