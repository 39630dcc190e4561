"""Device period of the two-frames-queued loop with (a) device-resident frames, (b) host frames copied in the
main stream, (c) host frames uploaded on the copy stream (evxgpu_encode_upload)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from cairo_b200 import gpu, synth
W, H, N, Q = 1920, 1080, 44, 16
host = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
for t in range(N): host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
L = gpu.lib()
L.evxgpu_encode_upload.argtypes = [C.c_void_p, C.c_void_p]
L.evxgpu_encode_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int]

def run(mode):
    p = gpu.Pipeline(W, H, 2, 0, 1)
    p.set_output(1)
    def sub(t):
        ft = 0 if t == 0 else 1
        if mode == 'device': rc = L.evxgpu_encode_submit(p.h, int(dev[t].data_ptr()), 1, ft, t, Q)
        elif mode == 'instream': rc = L.evxgpu_encode_submit(p.h, int(host[t].data_ptr()), 0, ft, t, Q)
        else:
            rc = L.evxgpu_encode_upload(p.h, int(host[t].data_ptr()))
            assert rc == 0
            rc = L.evxgpu_encode_submit(p.h, None, 0, ft, t, Q)
        assert rc == 0, rc
    for t in range(4):
        sub(t); p.encode_collect_bins()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sub(4)
    for t in range(5, N):
        sub(t); p.encode_collect_bins()
    p.encode_collect_bins()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    p.close()
    return 1e3 * dt / (N - 4)
for m in ['device', 'instream', 'upload', 'device', 'upload']:
    print(m, f"{run(m):.3f} ms/frame")
