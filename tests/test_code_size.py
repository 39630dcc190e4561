"""Round 2's central finding is that kernels sharing the device are limited by instruction supply: rolling the
search loops and calling the deblocking filter took the frame pipeline from 1 431 to 2 800 frames/s.  This guards the code
sizes that finding rests on (SASS bytes per kernel of the library that ships, from cuobjdump; no GPU needed)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cairo_b200", "libevxgpu.so")

# kernel name fragment -> upper bound in KB (measured at the end of round 2: 25 / 26.5 / 26.5 / 18.5 / 115 / 150 KB)
LIMITS = {
    "evx_inter_search": 32,
    "evx_search_follow": 34,
    "evx_deblock_follow": 34,
    "evx_deblock11": 24,
    "evx_wavefrontILi2": 135,
    "evx_wavefrontILi1": 175,
}


def _sizes():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is missing")
    sass = subprocess.run([cuobjdump, "-sass", SO], capture_output=True, text=True, check=True).stdout
    sizes, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            sizes[name] = 0
        elif name and re.match(r"\s+/\*[0-9a-f]+\*/\s", line):
            sizes[name] += 16
    return sizes


def test_hot_kernels_stay_small():
    sizes = _sizes()
    for frag, limit_kb in LIMITS.items():
        hits = {k: v for k, v in sizes.items() if frag in k}
        assert hits, f"kernel {frag} not found in the library"
        for k, v in hits.items():
            assert v <= limit_kb * 1024, f"{k}: {v / 1024:.1f} KB of SASS, limit {limit_kb} KB (see DESIGN 6a before unrolling anything)"


def test_library_is_sm_100a_only_and_uses_tma():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is missing")
    elf = subprocess.run([cuobjdump, "-lelf", SO], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", SO], capture_output=True, text=True, check=True).stdout
    assert "UTMALDG" in sass and "VIADDMNMX" in sass and "SYNCS" in sass
