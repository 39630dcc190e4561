"""Seeded, integer-only synthetic RGB sequences.

The same bytes feed the reference (oracle/_ref), the C oracle and the CUDA
path.  Everything is uint32/int64 numpy arithmetic, no floats, so the frames are
identical on every box.
"""
import numpy as np

_LCG_A = np.uint32(1664525)
_LCG_C = np.uint32(1013904223)


def _lcg_stream(seed: int, n: int) -> np.ndarray:
    """state_k for k=1..n of state = state*A + C (mod 2**32), vectorised."""
    a_pow = np.cumprod(np.full(n, _LCG_A, dtype=np.uint32), dtype=np.uint32)          # A^k, k=1..n
    geo = np.cumsum(np.concatenate(([np.uint32(1)], a_pow[:-1])), dtype=np.uint32)      # 1+A+..+A^(k-1)
    return a_pow * np.uint32(seed & 0xFFFFFFFF) + _LCG_C * geo


def frame(width: int, height: int, t: int, seed: int = 0, kind: str = "moving") -> np.ndarray:
    """One RGB8 frame, shape (height, width, 3), C-contiguous uint8.

    kind:
      moving  - moving gradient + moving 64x64 textured square + 2-bit noise on R
      static  - the t=0 'moving' frame without noise, repeated (-> all INTER_COPY)
      flat    - one constant colour
      noise   - full-range LCG noise on all channels
      dark    - 'moving' scaled to Y<32 (SAD-threshold tie rule, motion.cpp:138-140)
    """
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    if kind == "flat":
        out = np.empty((height, width, 3), dtype=np.uint8)
        out[..., 0] = (37 + seed) & 255
        out[..., 1] = (101 + 2 * seed) & 255
        out[..., 2] = (203 + 3 * seed) & 255
        return out
    if kind == "noise":
        s = _lcg_stream(1234 + seed * 1000 + t, width * height * 3)
        return ((s >> np.uint32(8)) & np.uint32(255)).astype(np.uint8).reshape(height, width, 3)
    tt = 0 if kind == "static" else t
    r = (((x + 2 * tt) * 255) // width) & 255
    g = (((y + tt) * 255) // height) & 255
    b = ((x + y + 3 * tt) >> 1) & 255
    r = np.broadcast_to(r, (height, width)).copy()
    g = np.broadcast_to(g, (height, width)).copy()
    b = b.copy()
    # textured square moving (5,3) px per frame
    sq = 64 if min(width, height) >= 128 else 16
    sx = (17 + 13 * seed + 5 * tt) % (width - sq)
    sy = (9 + 7 * seed + 3 * tt) % (height - sq)
    lx = np.arange(sq, dtype=np.int64)[None, :]
    ly = np.arange(sq, dtype=np.int64)[:, None]
    r[sy:sy + sq, sx:sx + sq] = (lx * 3 + ly) & 255
    g[sy:sy + sq, sx:sx + sq] = (ly * 5 + lx) & 255
    b[sy:sy + sq, sx:sx + sq] = (lx ^ ly) * 4 & 255
    if kind != "static":
        s = _lcg_stream(1234 + seed * 1000 + t, width * height)
        r = np.minimum(r + ((s >> np.uint32(8)) & np.uint32(3)).astype(np.int64).reshape(height, width), 255)
    out = np.stack([r, g, b], axis=-1)
    if kind == "dark":
        out = out >> 4
    return np.ascontiguousarray(out.astype(np.uint8))


def sequence(width: int, height: int, frames: int, seed: int = 0, kind: str = "moving") -> np.ndarray:
    return np.stack([frame(width, height, t, seed, kind) for t in range(frames)])
