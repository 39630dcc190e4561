#!/usr/bin/env python
"""bench.py -- 1080p P-frame encode throughput of the B200 pixel pipeline (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--extras 0|1]

A "step" is one P-frame of the configs[1] workload: 1080p synthetic sequence, quality 16, ring
of 2 slots (= 1 past reference frame), quarter-pel search, MPEG quantiser, deblocking on.  The
sequence's frame 0 (intra) is always part of the warm-up.  One process per GPU (torchrun for
N>1); each rank encodes its own independent stream (no collective on the data path -- nothing
reduces), so scaling is weak and `value` is the frames all ranks encoded / the slowest rank's
device time.

  value  : frames already resident in HBM -> pixel pipeline (K1 convert, search follower, wavefront rows,
           deblocking follower, K8 binarisation) -> the slice's bin string on the host (its D2H inside the timed
           region, host arithmetic coder excluded); the frames of the stream follow each other macroblock by
           macroblock on the device (ten frame slots).
  e2e    : evx1_encoder::submit/collect (the two halves of the reference's encode(), include/evx1_c.h) with HOST
           frames in pinned memory -> EVX1 bitstream bytes: H2D, kernels, D2H and the host arithmetic coder;
           e2e.synchronous is the same loop through evx1_encoder::encode, one frame at a time.
  --impl reference : the unmodified reference (oracle/_ref, built from /root/reference by
           oracle/Makefile) through the same public API on the host CPU.

HOW THE K STEPS ARE TIMED.  A window of K frames of a pipelined encoder is mostly fill and drain when K is 20 (a frame
is 2 ms on the device, a new one starts every 0.7 ms), and 17 ms of timed work is not a measurement.  So the stream is
run for `windows` consecutive windows of K frames (at least three, at least one second in total), between a barrier +
synchronize on both sides; the device time at which every frame's results had left the device is taken from CUDA events
on the stream that frame ran on (evxgpu_timeline_mark / evxgpu_last_done_ms), and `ms_per_step` is the MEDIAN window
divided by K.  `timing.first_window_ms` is the window that starts on an idle device (fill included), `timing.total_ms`
the whole region.  The end-to-end loop is timed the same way with the host clock at each window's last collect().

PARITY GATE.  Before anything is printed, rank 0 compares the bytes of the first frames of the timed encoders with the
bitstreams the UNMODIFIED reference produced for the same frames in this run (the cpu_baseline leg), and every rank
compares the pipelined and the synchronous encoders byte by byte over the first K frames.  A mismatch ends the run.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

# Independent video streams run on independent CUDA streams; with the default of 8 hardware work queues,
# 16 streams alias onto them and serialise falsely (measured: 16-stream pixel pipeline 2286 -> 2762 frames/s).
# Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, QUALITY, REF_COUNT, SEQ_FRAMES = 1920, 1080, 16, 2, 60
METRIC = "1080p P-frame encode frames/s per B200 (+1/2/4/8-GPU streams), bit-exact"
WORKLOAD = "configs[1]: 1080p synthetic 60-frame sequence, quality 16, 1 reference frame (ring of 2), quarter-pel ME"
# algorithmic integer ops of one full-pel candidate / one sub-pel test
OPS_FULLPEL, OPS_SUBPEL = 1024, 2560
LOOKAHEAD = int(os.environ.get("EVX_BENCH_LOOKAHEAD", "20"))      # frames between submit() and collect(): ten on the device + ten with the coder threads (16: 2 650, 20: 2 790 frames/s)


def config_block():
    """The same dict in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "streams_per_gpu": 1, "width": W, "height": H, "quality": QUALITY, "ref_count": REF_COUNT,
            "l2": f"every step reads a different 6.2 MB input frame ({SEQ_FRAMES} distinct frames = 373 MB > 126 MB L2; the sequence "
                  "wraps inside its P-frames), so no input is served from a previous step's L2 lines; the reference planes a "
                  "P-frame reads are the previous step's output by construction"}


def coder_threads_for(world):
    """Arithmetic-coder threads of a single-stream session (evx1_config::coder_threads): as many as this rank has host cores,
    between 2 and 6.  With four cores per GPU (the eight-GPU box) six coder threads plus the submitting thread thrash:
    measured with the process pinned to four cores, e2e 1 616 (six threads) / 1 464 (three) / 1 853 (four) frames/s."""
    env = int(os.environ.get("EVX_BENCH_CODER_THREADS", "0"))
    if env > 0:
        return env
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(2, min(6, cores // max(1, world)))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def synth_frames(w, h, n, seed, kind="moving"):
    """n frames of the seeded generator, made by a few threads (numpy releases the GIL in its inner loops)."""
    from cairo_b200 import synth
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda t: synth.frame(w, h, t, seed, kind), range(n)))


def streams_equal(a, abits, b, bbits, first):
    """Bit-exact comparison of two frames' streams; stream byte 7 (padding inside evx_header, uninitialised in the
    reference: SURVEY H7) is masked in the first frame."""
    if abits != bbits:
        return False
    n = (abits + 7) // 8
    x, y = np.array(a[:n], dtype=np.uint8), np.array(b[:n], dtype=np.uint8)
    if first and n > 7:
        x[7] = 0
        y[7] = 0
    if abits & 7 and n:
        m = (1 << (abits & 7)) - 1
        x[n - 1] &= m
        y[n - 1] &= m
    return bool((x == y).all())


def window_stats(done, K):
    """done[i]: time at which frame i of the timed region was finished.  Windows of K consecutive frames."""
    nwin = len(done) // K
    ends = [done[(i + 1) * K - 1] for i in range(nwin)]
    wins = [ends[0]] + [ends[i] - ends[i - 1] for i in range(1, nwin)]
    steady = wins[1:] if nwin > 1 else wins
    return {"windows": nwin, "median_ms": float(np.median(steady)), "first_window_ms": float(wins[0]),
            "min_ms": float(min(steady)), "max_ms": float(max(steady)), "total_ms": float(done[-1])}


def windows_for(K, est_ms_per_frame, min_ms):
    """How many windows of K frames: at least three, at least min_ms of timed work, the same on every rank."""
    return int(max(3, min(400, math.ceil(min_ms / max(1e-3, K * est_ms_per_frame)) + 1)))


# ---------------------------------------------------------------------------------- reference arm

def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import refharness as R
    variant = "r2"
    if not R.available(variant):
        # oracle/_ref travels with the snapshot; rebuild only where the reference sources exist
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not R.available(variant):
        emit({"impl": "reference", "unavailable": "oracle/_ref/libevxref_r2.so missing and /root/reference absent"})
        return 0
    n_streams = max(1, args.gpus)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    # the reference needs ~0.4 s per 1080p P-frame per core: K steps of one stream per GPU, every stream on a thread of its own
    frames = [synth_frames(W, H, min(warmup + steps, SEQ_FRAMES), s) for s in range(n_streams)]
    fidx = lambda t: t if t < SEQ_FRAMES else 1 + (t - 1) % (SEQ_FRAMES - 1)
    encs = []
    for s in range(n_streams):
        e = R.RefEncoder(variant)
        e.set_quality(QUALITY)
        encs.append(e)
    times = [0.0] * n_streams

    def work(s):
        for t in range(warmup):
            encs[s].encode(frames[s][fidx(t)])
        t0 = time.perf_counter()
        for t in range(warmup, warmup + steps):
            encs[s].encode(frames[s][fidx(t)])
        times[s] = time.perf_counter() - t0

    th = [threading.Thread(target=work, args=(s,)) for s in range(n_streams)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    elapsed = max(times)
    value = n_streams * steps / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * elapsed / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32",
        "data": "synthetic", "config": config_block(),
        "hardware": f"host CPU, {os.cpu_count()} logical cores; {n_streams} stream(s), one thread each (the reference encoder is single-threaded per stream)",
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": n_streams, "kind": "reference",
                         "sample": f"{n_streams} stream(s) x {steps} P-frames after {warmup} warm-up frames (frame 0 intra), one thread per stream, "
                                   "unmodified reference via evx1_encoder::encode, g++ -O2"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------- CPU baseline (reference) legs

def cpu_baseline_sample(frames, n):
    """The reference's CPU encoder on this box's host cores, bounded sample: the intra frame + n P-frames of the bench
    sequence.  Returns (cpu_baseline dict, [(bytes, bits)] of frames 0..n) -- the streams feed the parity gate."""
    import refharness as R
    if R.available("r2"):
        enc = R.RefEncoder("r2")
        enc.set_quality(QUALITY)
        streams = [enc.encode(frames[0])]
        t0 = time.perf_counter()
        for t in range(1, 1 + n):
            streams.append(enc.encode(frames[t]))
        dt = time.perf_counter() - t0
        return ({"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "reference",
                 "sample": f"{n} P-frames of the same 1080p sequence after the intra frame, single thread (the reference is single-threaded), oracle/_ref g++ -O2"},
                streams, "reference (oracle/_ref/libevxref_r2.so)")
    import oracleharness as O
    o = O.Oracle(W, H, REF_COUNT, 0, 1)
    streams = []
    t0 = 0.0
    for t in range(0, 1 + n):
        if t == 1:
            t0 = time.perf_counter()
        o.convert_in(frames[t]); o.encode_slice(0 if t == 0 else 1, t, QUALITY)
        streams.append(o.serialize())
        o.deblock(t)
    dt = time.perf_counter() - t0
    return ({"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
             "sample": f"{n} P-frames of the same 1080p sequence, single thread, oracle/evx_oracle.c gcc -O2"}, streams, "oracle port (slice payload only)")


def cpu_stage_split(frames, n=2):
    """Per-stage CPU times of the reference (encode.cpp:205-232: convert, encode_slice, serialize_slice, deblock) on n
    P-frames, single thread; and the same encoder on every host core at once (one stream per core)."""
    import refharness as R
    if not R.available("r2"):
        return None
    st = R.RefStage(W, H, "r2")
    acc = {"convert_image": 0.0, "encode_slice": 0.0, "serialize_slice": 0.0, "deblock": 0.0}
    for t in range(0, 1 + n):
        st.set_frame(0 if t == 0 else 1, t, QUALITY)
        t0 = time.perf_counter(); st.convert_in(frames[t])
        t1 = time.perf_counter(); st.encode_slice()
        t2 = time.perf_counter(); st.serialize()
        t3 = time.perf_counter(); st.deblock()
        t4 = time.perf_counter()
        if t:
            acc["convert_image"] += t1 - t0; acc["encode_slice"] += t2 - t1; acc["serialize_slice"] += t3 - t2; acc["deblock"] += t4 - t3
    out = {"ms_per_frame": {k: 1e3 * v / n for k, v in acc.items()}, "frames": n, "threads": 1}
    cores = os.cpu_count() or 1
    encs = []
    for _ in range(cores):
        e = R.RefEncoder("r2")
        e.set_quality(QUALITY)
        encs.append(e)
    gate = threading.Barrier(cores + 1)

    def work(i):
        encs[i].encode(frames[0])
        gate.wait()
        for t in range(1, 1 + n):
            encs[i].encode(frames[t])
        gate.wait()

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    for x in th:
        x.start()
    gate.wait()
    t0 = time.perf_counter()
    gate.wait()
    dt = time.perf_counter() - t0
    for x in th:
        x.join()
    out["all_cores"] = {"value": cores * n / dt, "unit": "frames/s", "cores": cores,
                        "sample": f"{cores} independent streams (one per logical core) x {n} P-frames, unmodified reference"}
    return out


# ---------------------------------------------------------------------------------- our arm: measurement loops

def device_run(gpu, dev, fidx, w, h, ring, linear, K, warmup, nwin, device, barrier, intra_every=0, frame_slots=0):
    """Frames resident in HBM -> submit / collect_bins through the C-ABI, as many frames in flight as the handle takes.
    -> window statistics from the device timeline, kernels launched, D2H bytes, total bins."""
    import torch
    pipe = gpu.Pipeline(w, h, ring, linear, 1, device=device, frame_slots=frame_slots)
    pipe.set_output(1)
    ftype = lambda t: 0 if t == 0 or (intra_every and t % intra_every == 0) else 1
    for t in range(warmup):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, QUALITY)
        pipe.encode_collect_bins()
    launches0 = pipe.launch_count()
    total = K * nwin
    barrier()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pipe.timeline_mark()
    done, d2h, bins = [], 0, 0
    cap = pipe.encode_capacity()
    ahead = min(cap - 1, total - 1)

    def take():
        nonlocal d2h, bins
        bins += pipe.encode_collect_bins()[1]
        d2h += pipe.d2h_bytes()
        done.append(pipe.last_done_ms())

    for t in range(warmup, warmup + ahead):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, QUALITY)
    for t in range(warmup + ahead, warmup + total):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, QUALITY)
        take()
    for _ in range(ahead):
        take()
    e1.record(stream)
    barrier()
    st = window_stats(done, K)
    st["bracket_ms"] = e0.elapsed_time(e1)
    res = {"stats": st, "launches": pipe.launch_count() - launches0, "d2h_bytes": d2h, "bins": bins, "frames": total, "slots": cap}
    pipe.close()
    return res


def e2e_run(enc, host, fidx, w, h, K, warmup, nwin, look, keep, barrier, intra_every=0, ptr_of=None):
    """evx1_encoder::submit/collect over K*nwin frames with `look` frames between a frame's submit and its collect.
    -> window statistics on the host clock, the first `keep` frames' streams, total bits."""
    import torch
    ptr_of = ptr_of or (lambda t: int(host[fidx(t)].data_ptr()))
    total = K * nwin
    look = max(1, min(look, total))
    coded, done, bits = [], [], 0
    barrier()
    t0 = time.perf_counter()

    def take():
        nonlocal bits
        d, b = enc.collect()
        done.append(1e3 * (time.perf_counter() - t0))
        bits += b
        if len(coded) < keep:
            coded.append((d.copy(), b))

    for t in range(warmup, warmup + total):
        if intra_every and t % intra_every == 0:
            enc.insert_intra()
        enc.submit((ptr_of(t), w, h))
        if t - warmup >= look:
            take()
    while len(done) < total:
        take()
    torch.cuda.synchronize()
    barrier()
    return {"stats": window_stats(done, K), "coded": coded, "bits": bits, "frames": total}


def kernel_pass(gpu, dev, fidx, w, h, ring, linear, warmup, n, device, intra_every=0):
    """Per-kernel times and work counters with the frames one after the other on the device (stand-alone kernels,
    frame_slots = 1): with frames overlapping, kernels of several frames share the SMs and a kernel's duration is not its own."""
    tp = gpu.Pipeline(w, h, ring, linear, 1, device=device, frame_slots=1)
    tp.enable_timing(True)
    tp.set_output(1)
    ftype = lambda t: 0 if t == 0 or (intra_every and t % intra_every == 0) else 1
    for t in range(warmup):
        tp.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, QUALITY)
        tp.encode_collect_bins()
    tp.counters(reset=True)
    tp.timing_sum(reset=True)
    for t in range(warmup, warmup + n):
        tp.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, QUALITY)
        tp.encode_collect_bins()
    ksum = {k: v / n for k, v in tp.timing_sum().items()}
    counts = [c / n for c in tp.counters_split()]
    tp.close()
    return ksum, counts


def search_roofline(ksum, counts, int_peak):
    """Integer roofline of the two search kernels from the per-frame kernel times and the candidates they evaluated."""
    k2_ms, k3_ms = ksum["inter_search"], ksum["wavefront"]
    ops_k2 = counts[0] * OPS_FULLPEL + counts[1] * OPS_SUBPEL
    ops_k3 = counts[2] * OPS_FULLPEL + counts[3] * OPS_SUBPEL
    a2 = ops_k2 / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else 0.0
    a3 = ops_k3 / (k3_ms * 1e-3) / 1e12 if k3_ms > 0 else 0.0
    return {"inter_search": {"ms": k2_ms, "ops": ops_k2, "achieved": a2, "frac": a2 / int_peak if int_peak else None},
            "wavefront": {"ms": k3_ms, "ops": ops_k3, "achieved": a3, "frac": a3 / int_peak if int_peak else None},
            "unit": "Tiop/s", "peak": int_peak}


def pinned_sequence(torch, w, h, n, seed):
    host = torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory()
    hnp = host.numpy()
    for t, f in enumerate(synth_frames(w, h, n, seed)):
        hnp[t] = f
    return host


def multi_stream_e2e(api, hosts, fidx, warmup, frames, device, frame_slots, look=3):
    """Aggregate frames/s of len(hosts) encoder sessions, each driven by a host thread of its own (hosts[i]: the stream's
    pinned frames)."""
    n_streams = len(hosts)
    encs = [api.evx1_encoder(device=device, ref_count=REF_COUNT, frame_slots=frame_slots, coder_threads=2) for _ in range(n_streams)]
    for e in encs:
        e.set_quality(QUALITY)
    gate = threading.Barrier(n_streams + 1)
    err = []

    def work(i):
        try:
            host = hosts[i]
            for t in range(warmup):
                encs[i].encode((int(host[fidx(t)].data_ptr()), W, H))
            gate.wait()
            n = 0
            for t in range(warmup, warmup + frames):
                encs[i].submit((int(host[fidx(t)].data_ptr()), W, H))
                if t - warmup >= look:
                    encs[i].collect(); n += 1
            while n < frames:
                encs[i].collect(); n += 1
        except Exception as ex:      # keep the barrier protocol alive, report after the join
            err.append(ex)
        gate.wait()

    th = [threading.Thread(target=work, args=(i,)) for i in range(n_streams)]
    for x in th:
        x.start()
    gate.wait()
    t0 = time.perf_counter()
    gate.wait()
    dt = time.perf_counter() - t0
    for x in th:
        x.join()
    del encs
    if err:
        raise err[0]
    return n_streams * frames / dt, dt


# ---------------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cairo_b200 import api, gpu, build, fanout
    # the libraries travel prebuilt; a rebuild (stale timestamps) is done by one rank only
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()

    K, warmup = max(1, args.steps), max(3, args.warmup)
    seed = rank
    host = pinned_sequence(torch, W, H, SEQ_FRAMES, seed)
    dev = host.to("cuda", non_blocking=False)
    fidx = lambda t: t if t < SEQ_FRAMES else 1 + (t - 1) % (SEQ_FRAMES - 1)       # wrap inside the P-frames
    frame_bytes = W * H * 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the reference on the host cores: the CPU baseline, and the bitstreams of the parity gate (rank 0)
    cpu_base, ref_streams, gate_against, stages = None, [], None, None
    if rank == 0:
        n_ref = 6 if world == 1 else 2
        try:
            cpu_base, ref_streams, gate_against = cpu_baseline_sample([host.numpy()[t] for t in range(1 + n_ref)], n_ref)
        except Exception as ex:
            raise SystemExit(f"bench.py: the reference leg of the parity gate failed: {ex}")
        if world == 1 and args.extras:
            try:
                stages = cpu_stage_split([host.numpy()[t] for t in range(3)], 2)
            except Exception as ex:
                stages = {"error": str(ex)}

    # ---- per-kernel times and work counters (stand-alone kernels, frame after frame)
    ksum, counts = kernel_pass(gpu, dev, fidx, W, H, REF_COUNT, 0, warmup, min(max(K, 8), 24), local_rank)

    # ---- value: device-resident frames through the pixel pipeline; what leaves the device per frame is the slice's bin
    # string, the input of the host arithmetic coder.  Nothing but submit/collect runs inside the timed region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    nwin = windows_for(K, 0.42, args.min_ms)
    dv = device_run(gpu, dev, fidx, W, H, REF_COUNT, 0, K, warmup, nwin, local_rank, barrier)

    # ---- e2e: the public API with host frames.  Two loops over the same frames: the reference's synchronous call
    # (encode), and its two halves (submit / collect) with LOOKAHEAD frames in between, so the host entropy stage of a frame
    # overlaps the device's work on the following ones -- the same bytes a few calls later.
    def fresh_encoder(**kw):
        kw.setdefault("coder_threads", coder_threads_for(world))
        e = api.evx1_encoder(device=local_rank, ref_count=REF_COUNT, **kw)
        e.set_quality(QUALITY)
        head = []
        for t in range(warmup):
            d, b = e.encode((int((dev if kw.get("device_frames") else host)[fidx(t)].data_ptr()), W, H))
            head.append((d.copy(), b))
        return e, head

    enc, head = fresh_encoder()
    sync_coded = []
    ent_ms = gpu_ms = 0.0
    d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for t in range(warmup, warmup + K):
        d, bits = enc.encode((int(host[fidx(t)].data_ptr()), W, H))
        sync_coded.append((d.copy(), bits))
        st = enc.stats()
        ent_ms += st["entropy_ms"]; gpu_ms += st["gpu_ms"]; d2h_bytes += st["d2h_bytes"]
    torch.cuda.synchronize()
    sync_ms = 1e3 * (time.perf_counter() - t0)
    del enc

    enc, head = fresh_encoder()
    nwin_e = windows_for(K, 0.45, args.min_ms)
    ee = e2e_run(enc, host, fidx, W, H, K, warmup, nwin_e, LOOKAHEAD, max(K, 240), barrier)      # (the first 240 streams feed the decode extra)
    clocks = sampler.stop()
    del enc

    # ---- the parity gate: bytes, not bit counts
    gate = {"frames": 0, "bytes_equal": True, "against": gate_against, "pipelined_vs_synchronous_frames": K}
    for i in range(K):
        if not streams_equal(ee["coded"][i][0], ee["coded"][i][1], sync_coded[i][0], sync_coded[i][1], False):
            raise SystemExit(f"bench.py: PARITY GATE: pipelined and synchronous streams differ in timed frame {i} (rank {rank})")
    if rank == 0:
        ours = head + ee["coded"]
        if gate_against and gate_against.startswith("reference"):
            for t, (rd, rb) in enumerate(ref_streams):
                if t >= len(ours):
                    break
                if not streams_equal(ours[t][0], ours[t][1], rd, rb, t == 0):
                    raise SystemExit(f"bench.py: PARITY GATE: frame {t} differs from the reference's bitstream ({ours[t][1]} vs {rb} bits)")
                gate["frames"] += 1
        else:      # the port's serialize() yields the slice payload without the 24/10-byte headers
            import oracleharness as O
            for t, (od, ob) in enumerate(ref_streams):
                if t >= len(ours):
                    break
                skip = (24 if t == 0 else 10) * 8
                d, b = ours[t]
                got = np.packbits(np.unpackbits(np.asarray(d, np.uint8), bitorder="little")[skip:b], bitorder="little")
                if not O.bits_equal(got, b - skip, od, ob):
                    raise SystemExit(f"bench.py: PARITY GATE: frame {t} differs from the oracle's slice")
                gate["frames"] += 1
    coded_all = ee["coded"]

    extras = {}
    if args.extras:
        extras = run_extras(args, torch, api, gpu, fanout, host, dev, fidx, head, coded_all, K, warmup, local_rank, rank, world, barrier)

    dev_ms_max, e2e_ms_max, sync_ms_max, dev_first_max = fanout.max_over_ranks(
        [dv["stats"]["median_ms"], ee["stats"]["median_ms"], sync_ms, dv["stats"]["first_window_ms"]], device="cuda")

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        hbm = float(peaks.get("hbm_gbs") or peaks.get("hbm_gbs_burst") or 6650.0)
        int_peak = gpu.lib().evxgpu_measure_int_peak(local_rank, 1)
        rf = search_roofline(ksum, counts, int_peak)
        traffic = None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f).get("evx_inter_search", {})
                traffic = tj.get("dram_bytes_read", 0) + tj.get("dram_bytes_write", 0)
                traffic_src = name
                break
        dominant = max(ksum, key=ksum.get)
        plane_bytes = 3 * W * H                      # int16 4:2:0 planes of a frame
        hbm_kernels = {}
        # The three streaming kernels.  `ms` is one launch per frame between its own events in the kernel pass (every launch reads
        # a frame that is not in L2; the event pair adds a few microseconds to a 6 us kernel, so the fraction is a lower bound);
        # `warm_ms` is the average of 50 back-to-back launches on the same 12 MB, which stay in L2: the kernel's own
        # duration without event overhead, NOT an HBM figure.
        b2b = {}
        try:
            sp = gpu.Pipeline(W, H, REF_COUNT, 0, 1, device=local_rank, frame_slots=1)
            for t in range(2):
                sp.encode(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, QUALITY)
            b2b = {"convert_in": sp.time_kernel(0, 50), "deblock": sp.time_kernel(1, 50), "convert_out": sp.time_kernel(2, 50)}
            sp.close()
        except Exception:
            b2b = {}
        kcold = dict(ksum)
        if extras.get("decode_kernels", {}).get("convert_out_ms"):
            kcold["convert_out"] = extras["decode_kernels"].pop("convert_out_ms")
        for name, key, nbytes in (("evx_rgb_to_yuv420_wide (K1)", "convert_in", frame_bytes + plane_bytes),
                                  ("evx_deblock (K4)", "deblock", 2 * plane_bytes + (W // 16) * ((H + 15) // 16) * 16),
                                  ("evx_yuv420_to_rgb_wide (K6)", "convert_out", frame_bytes + plane_bytes)):
            ms = kcold.get(key, 0.0)
            if ms > 0:
                gbs = nbytes / (ms * 1e-3) / 1e9
                hbm_kernels[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved_gbs": gbs, "frac": gbs / hbm,
                                     "warm_ms": b2b.get(key) if b2b.get(key, -1) > 0 else None}
        if extras.get("decode_kernels"):
            hbm_kernels.update(extras.pop("decode_kernels"))
        for v in hbm_kernels.values():
            v["frac"] = v["achieved_gbs"] / hbm
        line = {
            "metric": METRIC, "value": world * K / (dev_ms_max * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32",
            "data": "synthetic",
            "config": config_block(),
            "timing": {"method": "median over consecutive windows of `steps` frames of one continuous pipelined run between barrier+synchronize; "
                                 "per-frame completion times from CUDA events on the frame's own stream (value) / the host clock at collect() (e2e); max over ranks",
                       "value_windows": dv["stats"], "e2e_windows": ee["stats"],
                       "first_window_value": world * K / (dev_first_max * 1e-3),
                       "frames_timed": {"value": dv["frames"], "e2e": ee["frames"], "synchronous": K}},
            "scope": {"value": "frames resident in HBM -> K1 convert, search follower, wavefront rows, deblocking follower, K8 binarisation -> the slice's "
                               "bin string on the host (D2H inside the timed region, %d bytes per frame); host arithmetic coder excluded; %d frame slots: "
                               "consecutive frames follow each other macroblock by macroblock on the device" % (dv["d2h_bytes"] // dv["frames"], dv["slots"]),
                      "e2e": "evx1_encoder::submit/collect (the two halves of encode, %d frames between a frame's submit and its collect: ten on the device, "
                             "the rest with the coder threads), pinned host RGB -> EVX1 bitstream bytes: H2D, kernels, D2H of the bin string, host arithmetic "
                             "coder.  e2e.synchronous is the same through evx1_encoder::encode, one frame at a time" % LOOKAHEAD,
                      "kernels": "kernel_ms_per_step and the rooflines are from a separate pass with the frames one after the other (stand-alone kernels)"},
            "parity_gate": gate,
            "e2e": {"value": world * K / (e2e_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes,
                    "d2h_bytes_per_step": d2h_bytes // K,
                    "synchronous": {"value": world * K / (sync_ms_max * 1e-3), "unit": "frames/s", "api": "evx1_encoder::encode"},
                    "entropy_ms_per_step": ent_ms / K, "gpu_ms_per_step": gpu_ms / K, "coder_threads": coder_threads_for(world),
                    "bits_per_frame": ee["bits"] // ee["frames"]},
            "gpu_launches": int(round(dv["launches"] * K / dv["frames"])),
            "gpu_launches_note": "kernels of libevxgpu.so launched per window of `steps` frames in the value loop (%d over the %d frames timed)" % (dv["launches"], dv["frames"]),
            "kernel_ms_per_step": ksum,
            "roofline": {"bound": "int_alu", "kernel": "evx_inter_search (the motion-search kernel: all macroblocks x past references in parallel)",
                         "achieved": rf["inter_search"]["achieved"], "peak": int_peak, "unit": "Tiop/s", "frac": rf["inter_search"]["frac"], "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (profiles/%s); "
                                         "algorithmic bytes per launch = reference planes 6.27 MB + source planes 6.27 MB" % (traffic_src if traffic else "-"),
                         "peak_source": "evxgpu_measure_int_peak: dependency-free VIADDMNMX.S16x2 stream on all SMs, measured in this run "
                                        "(MEASURED_PEAKS.json has no integer figure)",
                         "algorithmic_ops_per_launch": rf["inter_search"]["ops"], "kernel_ms_per_launch": rf["inter_search"]["ms"],
                         "fullpel_candidates_per_launch": counts[0], "subpel_tests_per_launch": counts[1],
                         "ops_per_unit": {"fullpel_candidate": OPS_FULLPEL, "subpel_test": OPS_SUBPEL},
                         "dominant": {"kernel": "evx_wavefront (intra search + transform + reconstruction; raster-dependent, latency bound)" if dominant == "wavefront" else dominant,
                                      "ms_per_launch": ksum[dominant], "algorithmic_ops_per_launch": rf["wavefront"]["ops"],
                                      "achieved": rf["wavefront"]["achieved"], "frac": rf["wavefront"]["frac"], "critical_path_steps": W // 16 + 3 * ((H + 15) // 16 - 1)},
                         "pipeline": {"algorithmic_ops_per_frame": rf["inter_search"]["ops"] + rf["wavefront"]["ops"],
                                      "achieved": (rf["inter_search"]["ops"] + rf["wavefront"]["ops"]) / (dev_ms_max / K * 1e-3) / 1e12,
                                      "frac": (rf["inter_search"]["ops"] + rf["wavefront"]["ops"]) / (dev_ms_max / K * 1e-3) / 1e12 / int_peak if int_peak else None,
                                      "note": "both search kernels' ops over the pipelined frame period of `value`"},
                         "hbm_kernels": hbm_kernels,
                         "hbm_peak_gbs": hbm, "hbm_peak_kind": peaks_kind},
            "clocks": clocks,
        }
        line.update(extras)
        if cpu_base is not None and world == 1:
            line["cpu_baseline"] = cpu_base
            if stages:
                line["cpu_baseline"]["stages"] = stages
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_extras(args, torch, api, gpu, fanout, host, dev, fidx, head, coded, K, warmup, device, rank, world, barrier):
    """The other north_star configurations and API modes, each a short run of its own (named blocks of the JSON line)."""
    out = {}
    int_peak = gpu.lib().evxgpu_measure_int_peak(device, 1) if rank == 0 else 0.0
    frame_bytes = W * H * 3

    def agg(ms_per_frame):      # whole-job frames/s from every rank's frame period
        return world / (fanout.max_over_ranks([ms_per_frame], device="cuda")[0] * 1e-3)

    # ---- f3: device-resident frames through the public API (no 6.2 MB H2D per frame)
    try:
        e = api.evx1_encoder(device=device, ref_count=REF_COUNT, device_frames=True, coder_threads=coder_threads_for(world))
        e.set_quality(QUALITY)
        for t in range(warmup):
            e.encode((int(dev[fidx(t)].data_ptr()), W, H))
        r = e2e_run(e, None, fidx, W, H, K, warmup, windows_for(K, 0.45, args.min_ms / 2), LOOKAHEAD, min(K, 8), barrier, ptr_of=lambda t: int(dev[fidx(t)].data_ptr()))
        del e
        for i, (d, b) in enumerate(r["coded"]):
            if not streams_equal(d, b, coded[i][0], coded[i][1], False):
                raise RuntimeError(f"device-frame stream differs from the host-frame stream in frame {i}")
        out["e2e_device_frames"] = {"value": agg(r["stats"]["median_ms"] / K), "unit": "frames/s", "h2d_bytes_per_step": 0,
                                    "api": "evx1_encoder(device_frames) submit/collect: RGB frames already in HBM -> bitstream bytes on the host",
                                    "bytes_equal_to_host_frames": True, "windows": r["stats"]}
    except Exception as ex:
        out["e2e_device_frames"] = {"value": None, "error": str(ex)}

    # ---- decode: the same stream back to RGB ("plus decode fps"), bitstreams -> RGB in pinned memory
    try:
        rgb_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()

        def decode_run(pipelined):
            dec = api.evx1_decoder(device=device)
            for d, b in head:
                dec.decode(d, b, W, H, out=rgb_out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if pipelined:
                ahead = min(11, len(coded))
                for d, b in coded[:ahead]:
                    dec.submit(d, b)
                for d, b in coded[ahead:]:
                    dec.submit(d, b)
                    dec.collect(W, H, out=rgb_out)
                for _ in range(ahead):
                    dec.collect(W, H, out=rgb_out)
            else:
                for d, b in coded:
                    dec.decode(d, b, W, H, out=rgb_out)
            return 1e3 * (time.perf_counter() - t0) / len(coded)

        dec_sync, dec_pipe = decode_run(False), decode_run(True)
        out["decode"] = {"value": agg(dec_pipe), "unit": "frames/s",
                         "api": "evx1_decoder::submit/collect (twelve frames in flight, eight parser threads, the next frame handed to the device under the copy-out of the current one), bitstream -> RGB8 in pinned host memory",
                         "synchronous": agg(dec_sync), "frames": len(coded)}
        # K5 / K6 alone, for their HBM fractions
        if rank == 0:
            enc = gpu.Pipeline(W, H, REF_COUNT, 0, 1, device=device, frame_slots=1)
            dp = gpu.Pipeline(W, H, REF_COUNT, 0, 1, device=device, frame_slots=1)
            dp.enable_timing(True)
            n, acc = 0, {"convert_out": 0.0, "decode_recon": 0.0}
            for t in range(0, 6):
                tbl, rec = enc.encode(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, QUALITY)
                dp.decode(tbl, rec, 0 if t == 0 else 1, t)
                if t >= 2:
                    tm = dp.timing()
                    acc["convert_out"] += tm["convert_out"]; acc["decode_recon"] += tm["decode_recon"]
                    n += 1
            ks = {k: v / n for k, v in acc.items()}
            enc.close(); dp.close()
            plane_bytes = 3 * W * H
            dk = {}
            if ks.get("convert_out", 0) > 0:
                dk["convert_out_ms"] = ks["convert_out"]
            if ks.get("decode_recon", 0) > 0:
                nb = 3 * plane_bytes      # prediction read + coefficient records read (upper bound: every block coded) + reconstruction write
                dk["evx_decode_recon (K5, P-frame)"] = {"algorithmic_bytes": nb, "ms": ks["decode_recon"], "achieved_gbs": nb / (ks["decode_recon"] * 1e-3) / 1e9}
            out["decode_kernels"] = dk
    except Exception as ex:
        out["decode"] = {"value": None, "error": str(ex)}

    # ---- configs[2]: 1080p, ring of 4 (three past references), MPEG + adaptive QP + deblocking; and the linear quantiser
    for name, linear in (("r4", 0), ("r4_linear", 1)):
        try:
            ks, cn = kernel_pass(gpu, dev, fidx, W, H, 4, linear, warmup, 12, device)
            r = device_run(gpu, dev, fidx, W, H, 4, linear, K, warmup, windows_for(K, 0.62, args.min_ms / 2), device, barrier)
            blk = {"workload": "configs[2]: 1080p, quality 16, ring of 4 (3 past reference frames)%s, adaptive QP, deblocking" % (", linear quantiser" if linear else ", MPEG quantiser"),
                   "value": agg(r["stats"]["median_ms"] / K), "unit": "frames/s", "windows": r["stats"], "kernel_ms_per_step": ks,
                   "parity": "tests/test_gpu_fullsize.py::test_1080p_pipelined_stream_equals_reference[%s]" % name}
            if rank == 0:
                rf = search_roofline(ks, cn, int_peak)
                tr = None
                tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
                if os.path.exists(tp):
                    with open(tp) as f:
                        tj = json.load(f).get("evx_inter_search (ring of 4)", {})
                    tr = (tj.get("dram_bytes_read") or 0) + (tj.get("dram_bytes_write") or 0) or None
                blk["roofline"] = {"bound": "int_alu", "kernel": "evx_inter_search, 3 references", "achieved": rf["inter_search"]["achieved"], "peak": int_peak,
                                   "unit": "Tiop/s", "frac": rf["inter_search"]["frac"], "traffic": tr,
                                   "traffic_note": "DRAM bytes per launch, ncu --set full (profiles/r02_traffic.json); algorithmic: 4 plane sets x 6.27 MB", "dominant": {"kernel": "evx_wavefront", "ms_per_launch": ks["wavefront"],
                                                                                                       "achieved": rf["wavefront"]["achieved"], "frac": rf["wavefront"]["frac"]}}
            if not linear:
                e = api.evx1_encoder(device=device, ref_count=4, coder_threads=coder_threads_for(world))
                e.set_quality(QUALITY)
                for t in range(warmup):
                    e.encode((int(host[fidx(t)].data_ptr()), W, H))
                rr = e2e_run(e, host, fidx, W, H, K, warmup, windows_for(K, 0.65, args.min_ms / 2), LOOKAHEAD, 0, barrier)
                del e
                blk["e2e"] = {"value": agg(rr["stats"]["median_ms"] / K), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes, "windows": rr["stats"]}
            out[name] = blk
        except Exception as ex:
            out[name] = {"value": None, "error": str(ex)}

    # ---- configs[3]: 3840x2160, periodic intra every 30 frames
    try:
        w4, h4, n4 = 3840, 2160, 12
        host4 = pinned_sequence(torch, w4, h4, n4, rank)
        dev4 = host4.to("cuda")
        f4 = lambda t: t if t < n4 else 1 + (t - 1) % (n4 - 1)
        K4 = 30                                          # one intra period per window
        ks, cn = kernel_pass(gpu, dev4, f4, w4, h4, 2, 0, 3, 10, device)
        r = device_run(gpu, dev4, f4, w4, h4, 2, 0, K4, 3, 4, device, barrier, intra_every=30)
        e = api.evx1_encoder(device=device, ref_count=2, periodic_intra=30, coder_threads=coder_threads_for(world))
        e.set_quality(QUALITY)
        for t in range(3):
            e.encode((int(host4[f4(t)].data_ptr()), w4, h4))
        rr = e2e_run(e, host4, f4, w4, h4, K4, 3, 4, LOOKAHEAD, 0, barrier)
        del e
        blk = {"workload": "configs[3]: 3840x2160, quality 16, ring of 2, an intra frame every 30 frames (windows of 30 frames = one intra period)",
               "value": agg(r["stats"]["median_ms"] / K4), "unit": "frames/s", "windows": r["stats"], "kernel_ms_per_step_p_frames": ks,
               "e2e": {"value": agg(rr["stats"]["median_ms"] / K4), "unit": "frames/s", "h2d_bytes_per_step": w4 * h4 * 3, "windows": rr["stats"]},
               "parity": "tests/test_gpu_fullsize.py::test_4k_periodic_intra_stream_equals_reference"}
        if rank == 0:
            rf = search_roofline(ks, cn, int_peak)
            blk["roofline"] = {"bound": "int_alu", "kernel": "evx_inter_search at 4K", "achieved": rf["inter_search"]["achieved"], "peak": int_peak, "unit": "Tiop/s",
                               "frac": rf["inter_search"]["frac"], "dominant": {"kernel": "evx_wavefront", "ms_per_launch": ks["wavefront"],
                                                                               "achieved": rf["wavefront"]["achieved"], "frac": rf["wavefront"]["frac"]}}
        out["uhd_intra30"] = blk
        del dev4, host4
    except Exception as ex:
        out["uhd_intra30"] = {"value": None, "error": str(ex)}

    # ---- configs[4]: 64 independent 1080p streams over the ranks of this run (fanout.streams_of_rank), end to end
    try:
        from cairo_b200 import synth
        mine = fanout.streams_of_rank(64, rank, world)
        uniq, wu = 6, 2
        nfr = 24 * 64 // max(1, len(mine))              # the same number of timed frames per GPU whatever the partition
        # Frames in flight per stream: about 24 across the streams of a GPU saturate it (measured: 16 streams x 1 slot 2 958,
        # 8 x 3 2 889, 4 x 4 2 752, 2 x 8 2 604, 1 x 8 2 432 frames/s; 16 x 2 only 2 260) -- many streams fill the device
        # by themselves, frame after frame within each.
        slots = max(1, min(10, 24 // max(1, len(mine))))
        # distinct content per stream: four generator seeds, each shifted horizontally by 16 * (stream // 4) samples
        base = {s: synth_frames(W, H, uniq, 1000 + s) for s in sorted({m % 4 for m in mine})}
        hosts = [torch.empty((uniq, H, W, 3), dtype=torch.uint8).pin_memory() for _ in mine]

        def fill(k):
            m = mine[k]
            for t in range(uniq):
                hosts[k].numpy()[t] = np.roll(base[m % 4][t], 16 * (m // 4), axis=1)

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, range(len(mine))))
        wrap = lambda t: t if t < uniq else 1 + (t - 1) % (uniq - 1)
        barrier()
        fps, dt = multi_stream_e2e(api, hosts, wrap, wu, nfr, device, slots, look=max(3, slots + 3))
        total_frames = fanout.sum_over_ranks([len(mine) * nfr], device="cuda")[0]
        worst = fanout.max_over_ranks([dt], device="cuda")[0]
        out["fanout64"] = {"workload": "configs[4]: 64 independent 1080p streams (distinct content: 4 generator seeds x 16 horizontal shifts, 6 distinct frames each, "
                                       "wrapping inside the P-frames), quality 16, ring of 2, partitioned round-robin over the ranks (cairo_b200.fanout.streams_of_rank); "
                                       "one host thread + two coder threads per stream, evx1_encoder::submit/collect with pinned host frames -> bitstreams",
                           "value": total_frames / worst, "unit": "frames/s", "streams": 64, "streams_per_gpu": len(mine), "frames_per_stream": nfr,
                           "host_cores": os.cpu_count(), "frame_slots_per_stream": slots, "seconds": worst}
        del hosts
    except Exception as ex:
        out["fanout64"] = {"value": None, "error": str(ex)}
    return out


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else the process (NCCL's version banner,
    library chatter) writes to fd 1 has been pointed at stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=56)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--extras", type=int, default=1, help="0: only the headline configuration (configs[1]); 1: also decode, device frames, configs[2], [3], [4]")
    ap.add_argument("--min-ms", dest="min_ms", type=float, default=1000.0, help="timed work per headline loop (windows of --steps frames are repeated to reach it)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
