"""bench.py's measurement arithmetic and parity comparison (no GPU): window statistics, window count, the byte-level
stream comparison of the parity gate, and that both arms print the same `config` block."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("evx_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


def test_window_stats_median_excludes_the_fill_window():
    K = 4
    # frame completion times: the first window starts on an idle device (fill), then one frame every 0.5 ms, one hiccup
    done = [3.0, 3.5, 4.0, 4.5] + [4.5 + 0.5 * (i + 1) for i in range(8)] + [9.0, 9.5, 10.0, 12.5]
    st = bench.window_stats(done, K)
    assert st["windows"] == 4
    assert abs(st["first_window_ms"] - 4.5) < 1e-9
    assert abs(st["median_ms"] - 2.0) < 1e-9            # windows 2..4: 2.0, 2.0, 4.0
    assert abs(st["max_ms"] - 4.0) < 1e-9 and abs(st["min_ms"] - 2.0) < 1e-9
    assert abs(st["total_ms"] - 12.5) < 1e-9
    one = bench.window_stats([1.0, 2.0], 2)                # a single window is its own median
    assert one["windows"] == 1 and one["median_ms"] == 2.0


def test_windows_for_reaches_the_minimum_and_is_bounded():
    assert bench.windows_for(20, 0.4, 1000.0) >= 125        # 20 frames x 0.4 ms = 8 ms per window
    assert bench.windows_for(200, 0.4, 1000.0) >= 13
    assert bench.windows_for(10 ** 6, 1.0, 1000.0) == 3     # never fewer than three windows
    assert bench.windows_for(1, 0.001, 10 ** 9) == 400      # nor an unbounded number


def test_streams_equal_is_bit_exact_with_the_h7_mask():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, 40).astype(np.uint8)
    b = a.copy()
    assert bench.streams_equal(a, 317, b, 317, False)
    assert not bench.streams_equal(a, 317, b, 318, False)
    b[7] ^= 0x10                                             # padding byte inside evx_header: first frame only
    assert bench.streams_equal(a, 317, b, 317, True) and not bench.streams_equal(a, 317, b, 317, False)
    b = a.copy()
    b[39] ^= 0x40                                            # bit 318: beyond the 317 bits of the stream
    assert bench.streams_equal(a, 317, b, 317, False)
    b[39] ^= 0x01                                            # bit 312: inside
    assert not bench.streams_equal(a, 317, b, 317, False)
    b = a.copy()
    b[20] ^= 1
    assert not bench.streams_equal(a, 317, b, 317, True)


def test_both_arms_print_the_same_config_block():
    c = bench.config_block()
    assert c == bench.config_block()
    assert c["workload"].startswith("configs[1]") and c["streams_per_gpu"] == 1 and "l2" in c
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_block()') == 2       # the reference arm and ours
