#!/usr/bin/env python
"""Benchmark of the OETR hot path (feature-correlation transformer + overlap-box head) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 640|840] [--attention linear|full]
                    [--path hot|neck]

A step = one pass of the hot path over one batch of synthetic input: 32 pairs of 640x640 images, i.e. two
[32,256,20,20] fp32 feature maps (BASELINE.json configs[1]; --config 840: 16 pairs of 840x840, configs[3]); with N
GPUs every rank runs its own batch (weak scaling, replicated weights) and the boxes are gathered on every rank
(configs[2]).  One JSON line is printed by rank 0.

  value          pairs/s, inputs resident in HBM: CUDA-event time of exactly K steps, max over ranks; the K-step
                 region is repeated (--regions, default 15) and the MEDIAN region is reported
  e2e            the same metric through the C-ABI host-buffer entry points (oetr_forward_host_submit/_wait): pinned
                 host features -> H2D -> hot path -> D2H boxes (+ the gather when N > 1), every step, wall clock
  roofline       dominant kernel k_enc in the geometry that ships (sub-batch launches on several streams): per-tile wall
                 time measured on the device inside the timed region (globaltimer, one atomicAdd per tile) -> the
                 kernel's throughput with every SM running it = algorithmic FLOPs per tile x 148 / tile time, against
                 the measured fp16/bf16 tensor peak (MEASURED_PEAKS.json; burst figure: the region lasts milliseconds);
                 traffic = DRAM bytes per launch from the committed ncu capture
  cpu_baseline   the reference's algorithm with the reference's own arithmetic library (eager PyTorch fp32 on the host
                 cores, oracle/oetr_torch_eager.py -- pinned to the reference's outputs by tests/), bounded sample
  --impl reference   times that CPU implementation on the same workload as the reference arm (the reference itself is
                 pure Python under /root/reference, which does not exist on the GPU box)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-pairs/sec at 640x640"
# name -> (image side, feature-map side = side // 32, pairs per GPU, algorithmic GFLOP per pair linear / full attention)
CONFIGS = {"640": (640, 20, 32, 8.32, 10.94), "840": (840, 26, 16, 14.06, 21.55)}
N_ROTATE = 8                     # input batches cycled so that a step never finds its inputs in L2
N_SM = 148

# algorithmic FLOPs (multiply-add = 2), per token, of one k_enc launch = query phase of layer i + source phase of
# layer i+1: q, merge, k, v projections 4 * 2*256^2, MLP 2 * (2*256*512),
# Q.KV and K^T V 2 * (2*8*32*32).  The 3-term split executes 3x these on the tensor cores; only 1x is counted.
FLOPS_LAYER_PER_TOKEN = 4 * 2 * 256 * 256 + 2 * 2 * 256 * 512 + 2 * 2 * 8 * 32 * 32


def _traffic():
    """DRAM bytes of one k_enc launch (ncu --set full capture of the same code, profiles/)."""
    for name in ("r02_k_enc_metrics.json", "r01c_k_enc_metrics.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                m = json.load(f)
            return float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"]), name
        except Exception:
            continue
    return None, None


def _peaks(region_s):
    """Burst figure for a region of milliseconds, sustained for one of seconds (MEASURED_PEAKS.json)."""
    key = "bf16_tflops_sustained" if region_s >= 2.0 else "bf16_tflops"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p[key]), "measured (MEASURED_PEAKS.json, %s)" % key
    except Exception:
        return (1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)") if region_s >= 2.0 else \
               (1680.0, "fallback (B200_PROFILING.md: ~1.68 PFLOP/s burst)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (the timed
    region of 20 steps lasts ~25 ms, too short for `nvidia-smi -lms`), with nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop, self.t = None, None, False, None
        self.max_mhz = 0

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                return self.index
        return self.index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
        while not self.stop:
            try:
                mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append([str(mhz), str(self.max_mhz)] +
                                 ["Active" if mask & int(getattr(n, attr, 0)) else "Not Active" for _, attr in names])
            except Exception:
                break
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.nvml is not None:
            self.stop = True
            self.t.join(timeout=2)
            return
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(int(r[0]))
                mx = max(mx, int(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml, 2 ms period" if self.nvml is not None else "nvidia-smi -lms 100"}


def _all_host_threads():
    import torch
    n = os.cpu_count() or 1
    try:  # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core explicitly
        torch.set_num_threads(n)
    except Exception:
        pass
    return torch.get_num_threads()


def cpu_reference_pairs_per_s(cfg, n_pairs, reps, attention="linear"):
    """The reference's algorithm with the reference's own arithmetic library: eager PyTorch fp32 on the host cores
    (oracle/oetr_torch_eager.py, pinned to the reference's outputs by tests/test_oracle_golden.py) on `n_pairs` pairs of
    the bench workload.  Returns (pairs/s, seconds, threads)."""
    import torch
    from oetr_b200 import weights
    from oracle import oetr_torch_eager as ote
    side, fm, _, _, _ = CONFIGS[cfg]
    threads = _all_host_threads()
    W = ote.prepare(weights.synthetic_hot_path_weights(0), "cpu")
    f1 = torch.from_numpy(weights.synthetic_features(n_pairs, fm, fm, seed=21, tag="cpu1"))
    f2 = torch.from_numpy(weights.synthetic_features(n_pairs, fm, fm, seed=21, tag="cpu2"))
    ote.hot_path(W, f1[:2], f2[:2], (side, side), (side, side), attention=attention)      # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        ote.hot_path(W, f1, f2, (side, side), (side, side), attention=attention)
    dt = time.perf_counter() - t0
    return n_pairs * reps / dt, dt, threads


def _workload(cfg, attention):
    side, fm, pairs, _, _ = CONFIGS[cfg]
    return ("batch=%d %dx%d pairs per GPU: hot path (8-layer correlation transformer, %s attention, + decoder + overlap "
            "head) on two [%d,256,%d,%d] fp32 feature maps" % (pairs, side, side, attention, pairs, fm, fm))


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the SAME workload (whole batch per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    side, fm, pairs, _, _ = CONFIGS[args.config]
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_pairs_per_s(args.config, pairs, 1, args.attention)
    pps, dt, threads = cpu_reference_pairs_per_s(args.config, pairs, args.steps, args.attention)
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload(args.config, args.attention), "pairs_per_gpu": pairs,
                   "note": "CPU arm: one process, whole batch per step, eager PyTorch fp32 on %d threads" % threads},
        "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "%d pairs x %d steps; eager PyTorch fp32 restatement of the reference path "
                                   "(oracle/oetr_torch_eager.py: the reference's own operators, pinned to its outputs)" % (
                                       pairs, args.steps)},
        "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    # libraries (NCCL's version banner, ...) write to fd 1: keep stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import ctypes

    import torch
    import torch.distributed as dist

    import oetr_b200
    from oetr_b200 import cabi, weights
    from oetr_b200.distributed import BoxGather

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    side, fm, B, gflop_lin, gflop_full = CONFIGS[args.config]
    flops_per_pair = (gflop_full if args.attention == "full" else gflop_lin) * 1e9
    K, Wm, R = args.steps, args.warmup, max(1, args.regions)
    lib = cabi.load_library()

    Wts = weights.synthetic_hot_path_weights(0)
    # `in_flight` independent batches are kept in flight, each on its own CUDA stream and handle (every step is still
    # one complete stream-ordered forward of one batch): a sub-batch is a serial chain of ~25 launches, and with one
    # batch in flight the tail of that chain leaves SMs idle at every step boundary
    lanes = max(1, args.in_flight)
    hots = [oetr_b200.OverlapHotPath(Wts, attention=args.attention, precision=args.precision, device=dev) for _ in range(lanes)]
    for h_ in hots:
        h_.set_chunk_pairs(args.chunk_pairs)
    hot = hots[0]
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    # N_ROTATE different resident input batches (> 126 MB L2 together)
    base1 = weights.synthetic_features(B, fm, fm, seed=100 + rank, tag="b1")
    base2 = weights.synthetic_features(B, fm, fm, seed=100 + rank, tag="b2")
    feats = []
    for i in range(N_ROTATE):
        feats.append((torch.from_numpy(np.roll(base1, i, axis=0)).to(dev), torch.from_numpy(np.roll(base2, i, axis=0)).to(dev)))
    # boxes of every rank on every rank (configs[2]): one gather per step, issued on a dedicated communication stream
    # so that the lanes never wait for the rendezvous (oetr_b200/distributed.py)
    gathers = [BoxGather(B, dev) if world > 1 else None for _ in range(lanes)]
    hw = (side, side)

    def step(i, lane):
        f1, f2 = feats[i % N_ROTATE]
        with torch.cuda.stream(lane_streams[lane]):
            b1, b2 = hots[lane].forward(f1, f2, hw, hw, clamp=True)
            if world > 1:
                gathers[lane].submit(b1, b2)
        return b1, b2

    def sync_all():
        for g in gathers:
            if g is not None:
                g.wait()
        torch.cuda.synchronize()

    main = torch.cuda.current_stream(dev)
    for i in range(Wm * lanes):
        step(i, i % lanes)
    sync_all()
    for h_ in hots:
        h_.poll_error()
    lib.oetr_debug_cycles(None, -1, 1)                         # device-side per-tile timers on (one atomicAdd per tile)
    dbg = (ctypes.c_ulonglong * 48)()
    lib.oetr_debug_cycles(dbg, 48, 1)
    region_ms = []
    it = Wm * lanes
    with ClockSampler(local) as clk:
        for _ in range(R):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for st in lane_streams:
                st.wait_stream(main)
            for i in range(K):
                step(it + i, i % lanes)
            for lane, st in enumerate(lane_streams):
                if gathers[lane] is not None:
                    gathers[lane].join(st)                     # the region ends when the gathered boxes are there
                main.wait_stream(st)
            ev1.record()
            torch.cuda.synchronize()
            region_ms.append(ev0.elapsed_time(ev1))
            it += K
    n_dbg = lib.oetr_debug_cycles(dbg, 48, 1)
    lib.oetr_debug_cycles(None, -1, 0)
    launches = hot.last_launch_count * K
    for h_ in hots:
        h_.poll_error()
    t = torch.tensor(region_ms, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # every region: max over ranks
    ms = float(t.median().item())
    region_all = [float(v) for v in t.tolist()]

    # gathered boxes: this rank's slice is its own boxes, and every rank holds the same global tensor
    gather_check = None
    if world > 1:
        b1, b2 = step(it, 0)
        sync_all()
        g = gathers[0].result()
        own = torch.stack([b1, b2], dim=1)
        ok_own = bool(torch.equal(g[rank * B:(rank + 1) * B], own))
        cks = torch.tensor([float(g.double().sum().item()), float(g.double().abs().sum().item())], device=dev, dtype=torch.float64)
        lo, hi_ = cks.clone(), cks.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        gather_check = {"own_slice_equal": ok_own, "same_on_all_ranks": bool(torch.equal(lo, hi_)), "rows": int(g.shape[0])}
        assert ok_own and gather_check["same_on_all_ranks"], gather_check

    # e2e: host buffers through the C ABI, H2D + D2H (+ gather) inside the timed region
    h1 = torch.from_numpy(base1).pin_memory()
    h2 = torch.from_numpy(base2).pin_memory()
    hn1, hn2 = h1.numpy(), h2.numpy()
    for _ in range(12):     # every one of the handle's 4 request slots: eager run, graph capture, first replay
        hot.forward_host(hn1, hn2, hw, hw)
    depth = max(1, min(4, args.e2e_in_flight))
    hb = torch.empty(B, 2, 4).pin_memory()

    def e2e_region():
        tickets = []
        t0 = time.perf_counter()
        for i in range(K):
            tickets.append(hot.submit_host(hn1, hn2, hw, hw))
            if len(tickets) == depth:
                eb1, eb2 = hot.wait_host(tickets.pop(0))
                if world > 1:
                    gathers[0].submit_host(eb1, eb2, hb)
        while tickets:
            eb1, eb2 = hot.wait_host(tickets.pop(0))
            if world > 1:
                gathers[0].submit_host(eb1, eb2, hb)
        if world > 1:
            gathers[0].wait()
        return time.perf_counter() - t0

    e2e_s = []
    for _ in range(5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e2e_s.append(e2e_region())
    t0 = time.perf_counter()
    for _ in range(K):
        hot.forward_host(hn1, hn2, hw, hw)
    e2e_blocking_s = time.perf_counter() - t0
    te = torch.tensor(e2e_s, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_med = float(te.median().item())
    gather_mode = gathers[0].mode if world > 1 else None
    gather_note = gathers[0].fallback_reason if world > 1 else None
    for g_ in gathers:
        if g_ is not None:
            g_.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # parity spot check of the timed configuration against the CPU oracle (4 pairs, unclamped boxes)
    from oracle import oetr_oracle as orc
    npar = 4
    want = orc.hot_path(Wts, base1[:npar], base2[:npar], hw, hw, attention=args.attention, clamp=False)
    g1, g2 = hot.forward(feats[0][0][:npar], feats[0][1][:npar], hw, hw, clamp=False)
    torch.cuda.synchronize()
    perr = max(np.abs(g1.cpu().numpy() - want["box1_raw"]).max(), np.abs(g2.cpu().numpy() - want["box2_raw"]).max()) / side

    region_s = ms * 1e-3
    peak, peak_src = _peaks(region_s)
    v = [int(x) for x in dbg]
    roof = None
    whole_tflops = flops_per_pair * B * K / (ms * 1e-3) / 1e12
    if n_dbg and v[3] > 0 and args.precision == "fp16":
        tile_ns = v[4] / v[3]
        tile_cycles = v[0] / v[3]
        # algorithmic FLOPs of one k_enc tile (query phase of a layer + source phase of the next): tokens of the tile
        # that are not padding x FLOPs per token
        L = fm * fm
        Lp = (L + 15) // 16 * 16
        tiles_per_set = (B * Lp + 127) // 128
        tok_per_tile = B * L / tiles_per_set
        flops_tile = FLOPS_LAYER_PER_TOKEN * tok_per_tile
        ach = flops_tile * N_SM / (tile_ns * 1e-9) / 1e12
        traffic, traffic_src = _traffic()
        roof = {"bound": "tensor", "kernel": "k_enc", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "flops_per_launch": flops_tile * 2 * tiles_per_set, "tile_us": tile_ns * 1e-3, "tile_cycles": tile_cycles,
                "tiles_timed": v[3], "mma_lane_cycles": {"issuing": (v[0] - v[1] - v[2]) / v[3], "waiting_operand_image": v[1] / v[3],
                                                         "waiting_weights": v[2] / v[3]},
                "share_of_step": (v[4] * 1e-6 / (R * K)) / (N_SM * ms / K),
                "note": "algorithmic FLOPs (3-term split products count once); achieved = FLOPs of one 128-token tile x 148 SMs / "
                        "mean wall time of a tile, measured on the device for every tile of the encoder-layer launches inside the "
                        "timed region (sub-batch launches on several streams, as shipped); share_of_step = SM-time of those "
                        "launches / (148 x step time)",
                "whole_path_tflops": whole_tflops, "whole_path_frac": whole_tflops / peak}
    else:
        roof = {"bound": "tensor", "kernel": "whole hot path", "achieved": whole_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": whole_tflops / peak, "traffic": None, "peak_source": peak_src}
    cpu_pps, cpu_dt, cpu_threads = cpu_reference_pairs_per_s(args.config, B, 2, args.attention)
    eager = gpu_baselines(torch, dev, Wts, feats, hw, B, args.attention, hot)
    value = world * B * K / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": _workload(args.config, args.attention),
                   "pairs_per_gpu": B, "precision": args.precision, "attention": args.attention,
                   "sub_batches": ("%s pairs each, own stream" % (args.chunk_pairs if args.chunk_pairs > 0 else "automatic (~56 encoder tiles)")) if args.chunk_pairs else "off",
                   "batches_in_flight": "%d (one CUDA stream + handle each; every step = one complete forward)" % lanes,
                   "regions": "%d regions of %d steps, median reported (all: %s ms)" % (R, K, ", ".join("%.2f" % x for x in region_all)),
                   "l2": "inputs rotate over %d resident batches (%.0f MB > 126 MB L2)" % (
                       N_ROTATE, N_ROTATE * 2 * base1.nbytes / 1e6),
                   "parallelism": "batch shards, replicated weights, boxes gathered on every rank (dedicated comm stream)" if world > 1 else "1 GPU"},
        "clocks": clk.summary(),
        "e2e": {"value": world * B * K / e2e_med, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * base1.nbytes),
                "d2h_bytes_per_step": int(B * 8 * 4), "api": "oetr_forward_host_submit/_wait (C ABI, pinned host feature buffers, %d requests in flight)%s" % (
                    depth, "; boxes gathered on every rank each step" if world > 1 else ""),
                "regions_s": [float(x) for x in te.tolist()],
                "blocking_value": world * B * K / e2e_blocking_s,
                "blocking_api": "oetr_forward_host (one request at a time; rank-local time)"},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": {"value": cpu_pps, "unit": "pairs/s", "cores": cpu_threads, "kind": "port",
                         "sample": "%d pairs x 2 runs (%.1f s), eager PyTorch fp32 restatement of the reference path on the host "
                                   "cores (oracle/oetr_torch_eager.py: the reference's own operators)" % (B, cpu_dt)},
        "gpu_eager_baseline": eager,
        "parity": {"box_err_over_image_side": float(perr), "bar": 1e-3, "pairs_checked": npar},
    }
    if gather_check is not None:
        gather_check["mode"] = gather_mode + (" (peer stores over NVLink, CUDA IPC; no collective)" if gather_mode == "peer" else " (all_gather_into_tensor on a communication stream)")
        if gather_note:
            gather_check["peer_mode_unavailable"] = gather_note
        line["gather_check"] = gather_check
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_neck(args):
    """--path neck: the same contract for the neck (input_proj -> PatchMerging -> input_proj2 on the two
    image sets of a batch; NOT the headline metric).  value: backbone features resident in HBM, CUDA events, median of the
    regions; e2e: oetr_neck_forward fed from pinned host buffers every step (H2D of the 1024-channel features inside the
    timed region, D2H of the 256-channel output); roofline: the whole neck's algorithmic FLOPs against the measured tensor
    peak (k_neck_conv holds 76 % of the device time, profiles/r02_neck_launches.txt); cpu_baseline: the reference's own
    PyTorch modules' arithmetic (torch conv2d / layer_norm, fp32) on the host cores, bounded sample."""
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.nn.functional as F

    from oetr_b200 import weights
    from oetr_b200.neck import NeckB200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and rank > 0:
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    side, fm, B, _, _ = CONFIGS[args.config]
    h = w = side // 16
    n = 2 * B
    K, Wm, R = args.steps, args.warmup, max(1, args.regions)
    W = weights.synthetic_neck_weights(0)
    neck = NeckB200(W, device=dev)
    host = [torch.from_numpy(weights.synthetic_backbone_features(n, h, w, seed=60 + i)).pin_memory() for i in range(2)]
    xs = [t.to(dev) for t in host]                            # 2 x 420 MB at 640: a step never finds its input in L2
    out = torch.empty(n, 256, h // 2, w // 2, device=dev)
    out_host = torch.empty(out.shape, dtype=torch.float32).pin_memory()
    for i in range(Wm):
        neck.forward(xs[i & 1], out=out)
    torch.cuda.synchronize()
    regions = []
    with ClockSampler(0) as clk:
        for _ in range(R):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for i in range(K):
                neck.forward(xs[i & 1], out=out)
            ev1.record()
            torch.cuda.synchronize()
            regions.append(ev0.elapsed_time(ev1))
    ms = statistics.median(regions)
    staging = torch.empty_like(xs[0])
    e2e = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            staging.copy_(host[i & 1], non_blocking=True)
            neck.forward(staging, out=out)
            out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        e2e.append(time.perf_counter() - t0)
    e2e_s = statistics.median(e2e)
    t, o = h * w, (h // 2) * (w // 2)
    flops_img = 2.0 * (t * 1024 * 256 + o * 256 * (256 * 16 + 128 * 64 + 128 * 256) + o * 512 * 256)
    peak, peak_src = _peaks(ms * 1e-3)
    ach = flops_img * n * K / (ms * 1e-3) / 1e12
    # parity spot check against the oracle (2 images) and the CPU baseline (torch fp32 on the host cores, bounded sample)
    from oracle import neck_oracle as nk
    got = neck.forward(xs[0][:2].contiguous()).cpu().numpy()
    want = nk.neck(W, host[0][:2].numpy())
    perr = float(np.abs(got - want).max() / want.std())
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    xc = host[0][:8]

    def torch_neck(x):
        p = F.conv2d(x, Wt["input_proj.weight"], Wt["input_proj.bias"])
        p = F.layer_norm(p.permute(0, 2, 3, 1), (256,), Wt["patchmerging.norm.weight"], Wt["patchmerging.norm.bias"]).permute(0, 3, 1, 2)
        outs = [F.conv2d(p, Wt["patchmerging.reductions.%d.weight" % i], Wt["patchmerging.reductions.%d.bias" % i], stride=2,
                         padding=(k - 2) // 2) for i, k in enumerate((4, 8, 16))]
        return F.conv2d(torch.cat(outs, 1), Wt["input_proj2.weight"], Wt["input_proj2.bias"])
    with torch.no_grad():
        torch_neck(xc[:2])
        t0 = time.perf_counter()
        torch_neck(xc)
        cpu_dt = time.perf_counter() - t0
    line = {
        "metric": "neck image-pairs/sec at %dx%d" % (side, side), "value": B * K / (ms * 1e-3), "unit": "pairs/s", "n_gpus": 1,
        "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
        "config": {"workload": "batch=%d %dx%d pairs: neck (input_proj -> PatchMerging -> input_proj2) on [%d,1024,%d,%d] fp32 backbone "
                               "features -> [%d,256,%d,%d]" % (B, side, side, n, h, w, n, h // 2, w // 2),
                   "regions": "%d regions of %d steps, median reported" % (R, K),
                   "l2": "inputs alternate between two resident batches (%.0f MB each > 126 MB L2)" % (xs[0].numel() * 4 / 1e6)},
        "clocks": clk.summary(),
        "e2e": {"value": B * K / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(xs[0].numel() * 4), "d2h_bytes_per_step": int(out.numel() * 4),
                "api": "oetr_neck_forward (C ABI via NeckB200.forward) fed from pinned host buffers, one request at a time"},
        "gpu_launches": neck.last_launch_count * K,
        "roofline": {"bound": "tensor", "kernel": "whole neck (k_neck_proj + k_neck_conv + k_neck_out)", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                     "flops_per_step": flops_img * n, "note": "algorithmic FLOPs; single fp16 products, so the cap is 1.0"},
        "cpu_baseline": {"value": 4 / cpu_dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "8 images (4 pairs), %.1f s, torch conv2d / layer_norm fp32 on the host cores (the reference modules' "
                                   "arithmetic)" % cpu_dt},
        "parity": {"feature_err_over_std": perr, "bar": 4e-3, "images_checked": 2},
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def gpu_baselines(torch, dev, Wts, feats, hw, B, attention, hot):
    """The same path as eager PyTorch on this GPU (the bar SURVEY.md 2b names): fp32 with TF32 off, and TF32 on +
    CUDA-graph replay (no launch overhead).  Measurement aid only."""
    out = {}
    try:
        from oracle import oetr_torch_eager as ote
        Wt = ote.prepare(Wts, dev)
        f1, f2 = feats[0]

        def timed(fn, n):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return B * n / (e0.elapsed_time(e1) * 1e-3)

        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        t12 = [None]

        def eager():
            t12[0] = ote.hot_path(Wt, f1, f2, hw, hw, attention=attention)
        out["fp32_eager"] = {"value": timed(eager, 5), "unit": "pairs/s",
                             "kind": "eager PyTorch fp32 (TF32 off) restatement of the same path, same GPU, same batch, device-resident"}
        g1, g2 = hot.forward(f1, f2, hw, hw)
        torch.cuda.synchronize()
        out["fp32_eager"]["max_box_diff_px"] = float(max((g1 - t12[0][0]).abs().max().item(), (g2 - t12[0][1]).abs().max().item()))
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            for _ in range(3):
                ote.hot_path(Wt, f1, f2, hw, hw, attention=attention)
        torch.cuda.current_stream(dev).wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            gout = ote.hot_path(Wt, f1, f2, hw, hw, attention=attention)
        out["tf32_cuda_graph"] = {"value": timed(graph.replay, 10), "unit": "pairs/s",
                                  "kind": "the same eager PyTorch path with TF32 matmuls/convolutions, captured in ONE CUDA graph and replayed",
                                  "max_box_diff_px": float(max((g1 - gout[0]).abs().max().item(), (g2 - gout[1]).abs().max().item()))}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    except Exception as e:  # measurement aid only
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="640", choices=sorted(CONFIGS), help="640: batch 32 of 640x640 pairs (BASELINE configs[1]); 840: batch 16 of 840x840 (configs[3])")
    ap.add_argument("--attention", default="linear", choices=["linear", "full"])
    ap.add_argument("--regions", type=int, default=15, help="the K-step timed region is repeated this many times; the median is reported")
    ap.add_argument("--chunk-pairs", type=int, default=int(os.environ.get("OETR_CHUNK_PAIRS", "-1")),
                    help="pairs per concurrently scheduled sub-batch (0 = off, -1 = automatic: 8 at 640x640)")
    ap.add_argument("--in-flight", type=int, default=int(os.environ.get("OETR_IN_FLIGHT", "3")),
                    help="independent batches kept in flight in the device-resident timed region")
    ap.add_argument("--e2e-in-flight", type=int, default=4, help="host requests kept in flight in the e2e leg (1..4)")
    ap.add_argument("--precision", default=os.environ.get("OETR_BENCH_PRECISION", "fp16"), choices=["fp16", "fp32"])
    ap.add_argument("--path", default="hot", choices=["hot", "neck"], help="hot: the headline hot path (default); neck: the same "
                    "contract for the neck kernels (SURVEY 8(f1)), one GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.path == "neck":
        run_neck(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
