#!/usr/bin/env python
"""bench.py -- 1080p P-frame encode throughput of the B200 pixel pipeline (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one P-frame of the configs[1] workload: 1080p synthetic sequence, quality 16, ring
of 2 slots (= 1 past reference frame), quarter-pel search, MPEG quantiser, deblocking on.  The
sequence's frame 0 (intra) is always part of the warm-up.  One process per GPU (torchrun for
N>1); each rank encodes its own independent stream (no collective on the data path -- nothing
reduces), so scaling is weak and `value` is the frames all ranks encoded / the slowest rank's
device time.

  value  : frames already resident in HBM -> pixel pipeline (K1 convert, K2 inter search, K3 wavefront, K8
           binarisation, K4 deblocking) -> the slice's bin string on the host (its D2H inside the timed region,
           host arithmetic coder excluded); three frames in flight, overlapping on the device row by row.
  e2e    : evx1_encoder::submit/collect (the two halves of the reference's encode(), include/evx1_c.h) with HOST
           frames in pinned memory -> EVX1 bitstream bytes: H2D, kernels, D2H and the host arithmetic coder;
           e2e.synchronous is the same loop through evx1_encoder::encode, one frame at a time.
  --impl reference : the unmodified reference (oracle/_ref, built from /root/reference by
           oracle/Makefile) through the same public API on the host CPU.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

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
    from cairo_b200 import synth
    n_streams = max(1, args.gpus)
    steps, warmup = args.steps, max(1, args.warmup)
    # bounded sample: the reference needs ~1 s per 1080p P-frame per core
    steps = min(steps, 8)
    warmup = min(warmup, 2)
    frames = [[synth.frame(W, H, t, s, "moving") for t in range(warmup + steps)] for s in range(n_streams)]
    encs = []
    for s in range(n_streams):
        e = R.RefEncoder(variant)
        e.set_quality(QUALITY)
        encs.append(e)
    times = [0.0] * n_streams

    def work(s):
        for t in range(warmup):
            encs[s].encode(frames[s][t])
        t0 = time.perf_counter()
        for t in range(warmup, warmup + steps):
            encs[s].encode(frames[s][t])
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
        "data": "synthetic", "config": {"workload": WORKLOAD, "streams": n_streams, "hardware": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": n_streams, "kind": "reference",
                         "sample": f"{n_streams} stream(s) x {steps} P-frames after {warmup} warm-up frames (frame 0 intra), one thread per stream, "
                                   "unmodified reference via evx1_encoder::encode, g++ -O2"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------- our arm

def cpu_baseline_sample():
    """The reference's CPU encoder on this box's host cores, bounded sample (rank 0, N=1 only)."""
    import refharness as R
    from cairo_b200 import synth
    if R.available("r2"):
        enc = R.RefEncoder("r2")
        enc.set_quality(QUALITY)
        n = 6
        frames = [synth.frame(W, H, t, 0, "moving") for t in range(1 + n)]
        enc.encode(frames[0])
        t0 = time.perf_counter()
        for t in range(1, 1 + n):
            enc.encode(frames[t])
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "reference",
                "sample": f"{n} P-frames of the same 1080p sequence after the intra frame, single thread (the reference is single-threaded), oracle/_ref g++ -O2"}
    import oracleharness as O
    o = O.Oracle(W, H, REF_COUNT, 0, 1)
    n = 4
    frames = [synth.frame(W, H, t, 0, "moving") for t in range(1 + n)]
    o.convert_in(frames[0]); o.encode_slice(0, 0, QUALITY); o.serialize(); o.deblock(0)
    t0 = time.perf_counter()
    for t in range(1, 1 + n):
        o.convert_in(frames[t]); o.encode_slice(1, t, QUALITY); o.serialize(); o.deblock(t)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{n} P-frames of the same 1080p sequence, single thread, oracle/evx_oracle.c gcc -O2"}


def multi_stream_e2e(api, host, fidx, warmup, frames, n_streams, device):
    """Aggregate frames/s of n_streams encoder sessions driven from n_streams host threads."""
    encs = [api.evx1_encoder(device=device, ref_count=REF_COUNT) for _ in range(n_streams)]
    for e in encs:
        e.set_quality(QUALITY)
    gate = threading.Barrier(n_streams + 1)

    def work(i):
        for t in range(warmup):
            encs[i].encode((int(host[fidx(t)].data_ptr()), W, H))
        gate.wait()
        encs[i].submit((int(host[fidx(warmup)].data_ptr()), W, H))
        if frames > 1:
            encs[i].submit((int(host[fidx(warmup + 1)].data_ptr()), W, H))
        for t in range(warmup + 2, warmup + frames):
            encs[i].submit((int(host[fidx(t)].data_ptr()), W, H))
            encs[i].collect()
        for _ in range(min(2, frames)):
            encs[i].collect()
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
    return n_streams * frames / dt


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cairo_b200 import api, gpu, synth, build
    # the libraries travel prebuilt; a rebuild (stale timestamps) is done by one rank only
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()

    steps, warmup = args.steps, max(3, args.warmup)
    nframes = warmup + steps
    seed = rank
    # distinct frames, larger than L2 in total (60 x 6.2 MB = 373 MB > 126 MB L2)
    uniq = min(nframes, SEQ_FRAMES)
    host = torch.empty((uniq, H, W, 3), dtype=torch.uint8).pin_memory()
    hnp = host.numpy()
    for t in range(uniq):
        hnp[t] = synth.frame(W, H, t, seed, "moving")
    dev = host.to("cuda", non_blocking=False)
    fidx = lambda t: t if t < uniq else 1 + (t - 1) % (uniq - 1)       # wrap inside the P-frames if K > 59

    stream = torch.cuda.current_stream()
    frame_bytes = W * H * 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- per-kernel times and work counters: a pass with the frames one after the other on the device
    # (EVXGPU_FRAME_OVERLAP=0).  With consecutive frames overlapping, kernels of two frames share the SMs and a
    # kernel's own duration is no longer a property of the kernel; the roofline is quoted on the kernel running alone.
    os.environ["EVXGPU_FRAME_OVERLAP"] = "0"
    tp = gpu.Pipeline(W, H, REF_COUNT, 0, 1, device=local_rank)
    tp.enable_timing(True)
    tp.set_output(1)
    tsteps = min(steps, 24)
    for t in range(warmup):
        tp.encode_submit(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, QUALITY)
        tp.encode_collect_bins()
    tp.counters(reset=True)
    tp.timing_sum(reset=True)
    for t in range(warmup, warmup + tsteps):
        tp.encode_submit(int(dev[fidx(t)].data_ptr()), 1, t, QUALITY)
        tp.encode_collect_bins()
    ksum = {k: v * steps / tsteps for k, v in tp.timing_sum().items()}
    c_inter_full, c_inter_sub, c_intra_full, c_intra_sub = [c * steps / tsteps for c in tp.counters_split()]
    tp.close()
    os.environ.pop("EVXGPU_FRAME_OVERLAP", None)

    pipe = gpu.Pipeline(W, H, REF_COUNT, 0, 1, device=local_rank)
    # ---- value: device-resident frames through the pixel pipeline (K1 colour conversion, K2 inter search, K3
    # wavefront, K8 binarisation, K4 deblocking); what leaves the device per frame is the slice's bin string, the
    # input of the host arithmetic coder.  Nothing but submit/collect runs inside the timed region.  Two frames are in
    # flight and overlap on the device: frame t+1's search and wavefront follow frame t's row by row.
    pipe.set_output(1)
    for t in range(warmup):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, QUALITY)
        pipe.encode_collect_bins()
    launches0 = pipe.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    value_d2h = 0
    # the frames the handle holds (two, or three frame slots) are queued before the oldest one's bins are collected: the
    # device never waits for the host
    ahead = min(pipe.encode_capacity() - 1, steps - 1)
    for t in range(warmup, warmup + ahead):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), 1, t, QUALITY)
    for t in range(warmup + ahead, nframes):
        pipe.encode_submit(int(dev[fidx(t)].data_ptr()), 1, t, QUALITY)
        pipe.encode_collect_bins()
        value_d2h += pipe.d2h_bytes()
    for _ in range(ahead):
        pipe.encode_collect_bins()
        value_d2h += pipe.d2h_bytes()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = pipe.launch_count() - launches0
    pipe.close()

    # ---- e2e: the public API with host frames.  Two loops over the same frames: the reference's
    # synchronous call (encode), and its two halves (submit / collect) with one frame in flight, so the
    # host entropy stage of frame n overlaps the device's work on frame n+1 -- the same bytes one call later.
    def fresh_encoder():
        e = api.evx1_encoder(device=local_rank, ref_count=REF_COUNT)
        e.set_quality(QUALITY)
        for t in range(warmup):
            e.encode((int(host[fidx(t)].data_ptr()), W, H))
        return e

    enc = fresh_encoder()
    sync_bits = 0
    ent_ms = gpu_ms = 0.0
    d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for t in range(warmup, nframes):
        _, bits = enc.encode((int(host[fidx(t)].data_ptr()), W, H))
        sync_bits += bits
        st = enc.stats()
        ent_ms += st["entropy_ms"]; gpu_ms += st["gpu_ms"]; d2h_bytes += st["d2h_bytes"]
    torch.cuda.synchronize()
    sync_s = time.perf_counter() - t0
    del enc

    enc = fresh_encoder()
    out_bits = 0
    coded = []                                   # the K frames' bitstreams (a few KB each), for the decode extra
    barrier()
    t0 = time.perf_counter()
    # six frames of lookahead: while frame t is handed over, three earlier frames overlap on the device (three frame
    # slots) and the ones before them are being entropy-coded on the session's coder threads; collect() returns them in order
    look = min(int(os.environ.get("EVX_BENCH_LOOKAHEAD", "6")), steps)
    for t in range(warmup, warmup + look):
        enc.submit((int(host[fidx(t)].data_ptr()), W, H))
    for t in range(warmup + look, nframes):
        enc.submit((int(host[fidx(t)].data_ptr()), W, H))
        d, b = enc.collect()
        out_bits += b
        coded.append((d.copy(), b))
    for _ in range(look):
        d, b = enc.collect()
        out_bits += b
        coded.append((d.copy(), b))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if out_bits != sync_bits:
        raise SystemExit(f"bench.py: pipelined and synchronous streams differ ({out_bits} vs {sync_bits} bits)")
    clocks = sampler.stop()
    del enc

    # ---- extra: the decoder on the same stream ("plus decode fps"), bitstreams -> RGB in pinned memory
    decode_extra = None
    try:
        prefix = api.evx1_encoder(device=local_rank, ref_count=REF_COUNT)
        prefix.set_quality(QUALITY)
        head = []
        for t in range(warmup):
            d, b = prefix.encode((int(host[fidx(t)].data_ptr()), W, H))
            head.append((d.copy(), b))
        del prefix
        rgb_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()

        def decode_run(pipelined):
            dec = api.evx1_decoder(device=local_rank)
            for d, b in head:
                dec.decode(d, b, W, H, out=rgb_out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if pipelined:
                ahead = min(7, len(coded))
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
            return len(coded) / (time.perf_counter() - t0)

        dec_sync, dec_pipe = decode_run(False), decode_run(True)
        decode_extra = {"value": dec_pipe, "unit": "frames/s", "api": "evx1_decoder::submit/collect (eight frames in flight, six parser threads), bitstream -> RGB8 in pinned host memory",
                        "synchronous": dec_sync, "frames": len(coded)}
    except Exception as ex:
        decode_extra = {"value": None, "error": str(ex)}

    # ---- configs[4] in miniature: several independent streams sharing this GPU (one host thread,
    # one handle, one CUDA stream each); aggregate end-to-end throughput through the public API
    ms_streams = max(1, min(args.streams, (os.cpu_count() or 1) // max(1, world)))     # one host thread per stream
    ms_frames = min(24, steps)
    os.environ["EVXGPU_FRAME_OVERLAP"] = "0"      # many streams fill the device by themselves: frame after frame within each
    ms_fps = multi_stream_e2e(api, host, fidx, warmup, ms_frames, ms_streams, local_rank)
    os.environ.pop("EVXGPU_FRAME_OVERLAP", None)

    from cairo_b200 import fanout
    dev_ms_max, e2e_ms_max, sync_ms_max = fanout.max_over_ranks([dev_ms, e2e_s * 1e3, sync_s * 1e3], device="cuda")
    ms_total = fanout.sum_over_ranks([ms_fps], device="cuda")[0]

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        k2_ms = ksum["inter_search"] / steps
        k3_ms = ksum["wavefront"] / steps
        ops_k2 = (c_inter_full * OPS_FULLPEL + c_inter_sub * OPS_SUBPEL) / steps
        ops_k3 = (c_intra_full * OPS_FULLPEL + c_intra_sub * OPS_SUBPEL) / steps
        int_peak = gpu.lib().evxgpu_measure_int_peak(local_rank, 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f).get("evx_inter_search", {})
            traffic = tj.get("dram_bytes_read", 0) + tj.get("dram_bytes_write", 0)
        dominant = max(ksum, key=ksum.get)
        achieved = ops_k2 / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else 0.0
        achieved_all = (ops_k2 + ops_k3) / ((k2_ms + k3_ms) * 1e-3) / 1e12 if k2_ms + k3_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": world * steps / (dev_ms_max * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": 1, "value_scope": "frames resident in HBM -> K1 convert, K2 inter search, K3 wavefront, K8 binarisation, K4 deblocking -> the slice's bin string "
                                      "on the host (D2H inside the timed region, %d bytes per frame); host arithmetic coder excluded; up to three frames in "
                                      "flight (three frame slots), consecutive frames overlap on the device row by row; kernel_ms_per_step and the roofline are from a "
                                      "separate pass with the frames one after the other (a kernel's duration next to another frame's kernels is "
                                      "not its own)" % (value_d2h // max(1, steps)),
                       "e2e_scope": "evx1_encoder::submit/collect (the two halves of encode, six frames of lookahead: three overlapping on the device, the rest on the coder threads), pinned host RGB -> EVX1 bitstream bytes: "
                                    "H2D, K1..K4 + device binarisation K8, D2H of the bin string, host arithmetic coder; all K bitstreams are on the host "
                                    "when the clock stops.  e2e.synchronous is the same through evx1_encoder::encode, one frame at a time",
                       "l2": f"every timed step reads a different 6.2 MB input frame ({uniq} distinct frames, {uniq * frame_bytes // 1000000} MB, resident in HBM; "
                             f"the sequence wraps only after 59 P-frames = 367 MB > 126 MB L2), so no input is served from a previous "
                             f"step's L2 lines; the reference planes a P-frame reads are the previous step's output by construction"},
            "e2e": {"value": world * steps / (e2e_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes,
                    "d2h_bytes_per_step": d2h_bytes // steps,
                    "synchronous": {"value": world * steps / (sync_ms_max * 1e-3), "unit": "frames/s", "api": "evx1_encoder::encode"},
                    "entropy_ms_per_step": ent_ms / steps, "gpu_ms_per_step": gpu_ms / steps,
                    "bits_per_frame": out_bits // steps},
            "gpu_launches": int(launches),
            "multi_stream": {"workload": "configs[4] in miniature: independent 1080p streams of the same content per GPU, one host thread each, "
                                         "evx1_encoder::submit/collect end to end (host frames -> bitstreams)",
                             "streams_per_gpu": ms_streams, "value": ms_total, "unit": "frames/s", "frames_per_stream": ms_frames,
                             "host_cores": os.cpu_count()},
            "decode": decode_extra,
            "kernel_ms_per_step": {k: v / steps for k, v in ksum.items()},
            "roofline": {"bound": "int_alu", "kernel": "evx_inter_search (the motion-search kernel: all macroblocks x past references in parallel)",
                         "achieved": achieved, "peak": int_peak, "unit": "Tiop/s", "frac": achieved / int_peak if int_peak > 0 else None, "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (profiles/r01_traffic.json); "
                                         "algorithmic bytes per launch = reference planes 6.27 MB + source planes 6.27 MB",
                         "peak_source": "evxgpu_measure_int_peak: dependency-free VIADDMNMX.S16x2 stream on all SMs, measured in this run "
                                        "(MEASURED_PEAKS.json has no integer figure)",
                         "algorithmic_ops_per_launch": ops_k2, "kernel_ms_per_launch": k2_ms,
                         "fullpel_candidates_per_launch": c_inter_full / steps, "subpel_tests_per_launch": c_inter_sub / steps,
                         "ops_per_unit": {"fullpel_candidate": OPS_FULLPEL, "subpel_test": OPS_SUBPEL},
                         "serial_kernel": {"kernel": "evx_wavefront (intra search + transform + reconstruction, raster-dependent: latency bound)",
                                           "algorithmic_ops_per_launch": ops_k3, "kernel_ms_per_launch": k3_ms,
                                           "achieved": ops_k3 / (k3_ms * 1e-3) / 1e12 if k3_ms > 0 else 0.0,
                                           "critical_path_steps": 120 + 3 * 67},
                         "search_kernels_combined_achieved": achieved_all,
                         "dominant_kernel_by_time": dominant,
                         "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_kind": peaks_kind},
            "clocks": clocks,
        }
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_sample()
            except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--streams", type=int, default=16, help="streams per GPU of the extra multi_stream measurement (capped at the host core count)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
