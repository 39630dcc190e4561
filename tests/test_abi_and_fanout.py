"""CPU-side checks: the shared libraries load and export every symbol the headers declare (no
compute call is made), and the N>1 host logic works on a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_headers_declare_a_sane_number_of_entry_points():
    assert len(_declared("evxgpu.h", "evxgpu_")) >= 25
    assert len(_declared("evx1_c.h", "evx1c_")) >= 15


def test_gpu_library_exports_every_declared_symbol():
    from cairo_b200 import build
    build.build_all()
    lib = ctypes.CDLL(os.path.join(ROOT, "cairo_b200", "libevxgpu.so"))
    for name in _declared("evxgpu.h", "evxgpu_"):
        assert hasattr(lib, name), name


def test_host_library_exports_every_declared_symbol():
    from cairo_b200 import api
    lib = api.lib()
    for name in _declared("evx1_c.h", "evx1c_"):
        assert hasattr(lib, name), name


def test_shipped_sass_has_no_dropped_packed_add_immediate():
    """build.lint_sass: the ptxas miscompile of the packed negation (VIADD.16x2 with a zero immediate)
    must not be present in the library that ships."""
    from cairo_b200 import build
    build.build_all()
    build.lint_sass(build.GPU_SO)


def test_no_device_means_loud_failure_not_fallback():
    """Without a GPU the product must refuse to run (status 5), never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cairo_b200 import gpu
    with pytest.raises(RuntimeError, match="status 5"):
        gpu.Pipeline(64, 64)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cairo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracleharness" not in text and "evx_oracle" not in text and "libevxref" not in text, os.path.join(dirpath, f)


def test_record_scatter_gather_round_trip():
    """include/evxgpu_records.h (through libevx1.so): records of the non-copy macroblocks -> persistent coefficient planes
    -> records; copy blocks leave the planes' stale values alone; same result as the numpy restatement."""
    import numpy as np
    from cairo_b200 import api, gpu
    aw, ah = 80, 48
    n = (aw // 16) * (ah // 16)
    rng = np.random.default_rng(5)
    tbl = np.zeros(n, dtype=gpu.BLOCK_DESC_DTYPE)
    tbl["block_type"] = rng.integers(0, 8, n)
    k = int((tbl["block_type"] & 4 == 0).sum())
    rec = rng.integers(-300, 300, (k, 384)).astype(np.int16)
    planes = [np.full((ah, aw), 7, np.int16), np.full((ah // 2, aw // 2), 8, np.int16), np.full((ah // 2, aw // 2), 9, np.int16)]
    want = [a.copy() for a in planes]
    gpu.records_to_planes(tbl, rec, want, aw, ah)
    assert api.scatter_records(tbl, rec, planes, aw, ah) == k
    assert all((a == b).all() for a, b in zip(planes, want))
    assert (planes[0] == 7).any() or k == n                    # the copy blocks' samples were not touched
    back = api.gather_records(tbl, planes, aw, ah)
    assert back.shape == rec.shape and (back == rec).all()


def test_stream_partition():
    from cairo_b200 import fanout
    assert fanout.streams_of_rank(64, 3, 8) == [3, 11, 19, 27, 35, 43, 51, 59]
    allocated = sorted(s for r in range(4) for s in fanout.streams_of_rank(10, r, 4))
    assert allocated == list(range(10))
    with pytest.raises(ValueError):
        fanout.streams_of_rank(4, 4, 4)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from cairo_b200 import fanout
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
mine = fanout.streams_of_rank(5, rank, 2)
fps = fanout.aggregate_throughput(10 * len(mine), 1000.0 * (rank + 1))
assert abs(fps - 50 / 2.0) < 1e-9, fps            # 50 frames over the slower rank's 2 s
mx = fanout.max_over_ranks([float(rank), 7.0 - rank])
assert mx == [1.0, 7.0], mx
dist.barrier()
dist.destroy_process_group()
print("ok", rank, mine)
"""


def test_two_rank_gloo_aggregation(tmp_path):
    port = 29000 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "ok 0 [0, 2, 4]" in outs[0] and "ok 1 [1, 3]" in outs[1]
