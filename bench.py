#!/usr/bin/env python
"""Benchmark of the OETR hot path (feature-correlation transformer + overlap-box head) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one pass of the hot path over one batch of synthetic input: 32 pairs of 640x640 images, i.e. two
[32,256,20,20] fp32 feature maps (BASELINE.json configs[1]); with N GPUs every rank runs its own batch (weak
scaling, replicated weights) and the boxes are all-gathered (configs[2]).  One JSON line is printed by rank 0.

  value          pairs/s, inputs resident in HBM, CUDA-event time of exactly K steps, max over ranks
  e2e            the same metric through the C-ABI host-buffer entry point (oetr_forward_host): pinned host
                 features -> H2D -> hot path -> D2H boxes, every step, wall clock
  roofline       dominant kernel (k_enc, one launch per encoder layer): algorithmic FLOPs / launch over the
                 CUDA-event launch duration measured inside the timed region, against the measured bf16/fp16
                 tensor peak (MEASURED_PEAKS.json, sustained figure: the kernel is timed inside a long step);
                 traffic = DRAM bytes per launch from the committed ncu capture (profiles/r01c_k_enc_metrics.json)
  cpu_baseline   the numpy port of the reference algorithm (oracle/) on the host cores, bounded sample
  --impl reference   times that same CPU implementation as the reference arm (the reference itself is pure
                 Python under /root/reference, which does not exist on the GPU box; the port is pinned to it by
                 tests/golden)
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
FEAT_HW = 20                     # 640 / 32
IMG = 640
PAIRS_PER_GPU = 32
N_ROTATE = 8                     # input batches cycled so that a step never finds its inputs in L2

# algorithmic FLOPs (multiply-add = 2), per token, of one k_enc launch = query phase of layer i + source phase of
# layer i+1: q, merge, k, v projections 4 * 2*256^2, MLP 2 * (2*256*512),
# Q.KV and K^T V 2 * (2*8*32*32).  The 3-term split executes 3x these on the tensor cores; only 1x is counted.
FLOPS_LAYER_PER_TOKEN = 4 * 2 * 256 * 256 + 2 * 2 * 256 * 512 + 2 * 2 * 8 * 32 * 32
# whole hot path per pair at L=400: 8.32 GFLOP
FLOPS_PER_PAIR = 8.32e9


def _traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "r01c_k_enc_metrics.json")) as f:
            m = json.load(f)
        return float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])
    except Exception:
        return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


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


def cpu_port_pairs_per_s(n_pairs, reps):
    """Time the numpy port (fp32, all BLAS threads) on `n_pairs` pairs of the bench workload."""
    from oetr_b200 import weights
    from oracle import oetr_oracle as orc
    try:  # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core explicitly
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    W = weights.synthetic_hot_path_weights(0)
    f1 = weights.synthetic_features(n_pairs, FEAT_HW, FEAT_HW, seed=21, tag="cpu1")
    f2 = weights.synthetic_features(n_pairs, FEAT_HW, FEAT_HW, seed=21, tag="cpu2")
    orc.hot_path(W, f1[:1], f2[:1], (IMG, IMG), (IMG, IMG), dtype=np.float32)      # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.hot_path(W, f1, f2, (IMG, IMG), (IMG, IMG), dtype=np.float32)
    dt = time.perf_counter() - t0
    return n_pairs * reps / dt, dt


def _host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the CPU implementation of the path, one bounded sample (2 pairs) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 2
    for _ in range(max(args.warmup, 1) - 1):
        cpu_port_pairs_per_s(n, 1)
    pps, dt = cpu_port_pairs_per_s(n, args.steps)
    cores = _host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch=32 640x640 pairs (hot path on [*,256,20,20] feature maps); CPU sample of "
                               "%d pairs per step" % n},
        "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs x %d steps, numpy fp32 port of the reference algorithm" % (n, args.steps)},
        "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    # libraries (NCCL's version banner, ...) write to fd 1: keep stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    import oetr_b200
    from oetr_b200 import weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = PAIRS_PER_GPU, args.steps, args.warmup

    Wts = weights.synthetic_hot_path_weights(0)
    # `in_flight` independent batches are kept in flight, each on its own CUDA stream and handle (every step is still
    # one complete stream-ordered forward of one batch): a sub-batch is a serial chain of ~25 launches, and with one
    # batch in flight the tail of that chain leaves SMs idle at every step boundary
    lanes = max(1, args.in_flight)
    hots = [oetr_b200.OverlapHotPath(Wts, attention="linear", precision=args.precision, device=dev) for _ in range(lanes)]
    for h_ in hots:
        h_.set_chunk_pairs(args.chunk_pairs)
    hot = hots[0]
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    # N_ROTATE different resident input batches (8 x 26 MB > 126 MB L2)
    base1 = weights.synthetic_features(B, FEAT_HW, FEAT_HW, seed=100 + rank, tag="b1")
    base2 = weights.synthetic_features(B, FEAT_HW, FEAT_HW, seed=100 + rank, tag="b2")
    feats = []
    for i in range(N_ROTATE):
        feats.append((torch.from_numpy(np.roll(base1, i, axis=0)).to(dev), torch.from_numpy(np.roll(base2, i, axis=0)).to(dev)))
    gathered = [torch.empty(world * B, 2, 4, device=dev) if world > 1 else None for _ in range(lanes)]

    def step(i, lane=None):
        f1, f2 = feats[i % N_ROTATE]
        if lane is None:
            b1, b2 = hot.forward(f1, f2, (IMG, IMG), (IMG, IMG), clamp=True)
            if world > 1:
                dist.all_gather_into_tensor(gathered[0], torch.stack([b1, b2], dim=1))
            return b1, b2
        with torch.cuda.stream(lane_streams[lane]):
            b1, b2 = hots[lane].forward(f1, f2, (IMG, IMG), (IMG, IMG), clamp=True)
            if world > 1:
                dist.all_gather_into_tensor(gathered[lane], torch.stack([b1, b2], dim=1))
        return b1, b2

    main = torch.cuda.current_stream(dev)
    for i in range(Wm * lanes):
        step(i, i % lanes)
    torch.cuda.synchronize()
    for h_ in hots:
        h_.poll_error()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        ev0.record()
        for st in lane_streams:
            st.wait_stream(main)
        for i in range(K):
            step(Wm + i, i % lanes)
        for st in lane_streams:
            main.wait_stream(st)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = hot.last_launch_count * K
    for h_ in hots:
        h_.poll_error()
    # roofline leg: the same K steps with the kernel profiler on.  It brackets every k_enc launch with CUDA events on
    # the launching stream, which needs the launches serialised on ONE stream, so sub-batch scheduling is off here:
    # the kernel is timed whole-batch (256 tiles), in isolation, inside a long step.
    hot.profile(True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(K):
        step(Wm + K + i)
    ev3.record()
    torch.cuda.synchronize()
    layer_ms, n_layer = hot.profile_read()
    serial_ms = ev2.elapsed_time(ev3)
    hot.profile(False)
    hot.poll_error()

    # e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    h1 = torch.from_numpy(base1).pin_memory()
    h2 = torch.from_numpy(base2).pin_memory()
    for _ in range(12):     # every one of the handle's 4 request slots: eager run, graph capture, first replay
        hot.forward_host(h1.numpy(), h2.numpy(), (IMG, IMG), (IMG, IMG))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # K requests, `depth` in flight (oetr_forward_host_submit / _wait): every step's features cross PCIe from pinned host
    # memory and every step's boxes are read back on the host, all inside the timed region
    hn1, hn2 = h1.numpy(), h2.numpy()
    depth = max(1, min(4, args.e2e_in_flight))
    t0 = time.perf_counter()
    tickets = []
    for i in range(K):
        tickets.append(hot.submit_host(hn1, hn2, (IMG, IMG), (IMG, IMG)))
        if len(tickets) == depth:
            eb1, eb2 = hot.wait_host(tickets.pop(0))
    while tickets:
        eb1, eb2 = hot.wait_host(tickets.pop(0))
    e2e_s = time.perf_counter() - t0
    # the same through the blocking call (one request at a time), for reference
    t0 = time.perf_counter()
    for _ in range(K):
        hot.forward_host(hn1, hn2, (IMG, IMG), (IMG, IMG))
    e2e_blocking_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t[0].item(), t[1].item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # parity spot check of the timed configuration against the CPU port (2 pairs)
    from oracle import oetr_oracle as orc
    want = orc.hot_path(Wts, base1[:2], base2[:2], (IMG, IMG), (IMG, IMG), clamp=False)
    g1, g2 = hot.forward(feats[0][0][:2], feats[0][1][:2], (IMG, IMG), (IMG, IMG), clamp=False)
    torch.cuda.synchronize()
    perr = max(np.abs(g1.cpu().numpy() - want["box1_raw"]).max(), np.abs(g2.cpu().numpy() - want["box2_raw"]).max()) / IMG

    peak, peak_src = _peaks()
    tokens = B * 2 * FEAT_HW * FEAT_HW
    roof = None
    if n_layer:
        ach = FLOPS_LAYER_PER_TOKEN * tokens / (layer_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "k_enc", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": _traffic(), "peak_source": peak_src,
                "flops_per_launch": FLOPS_LAYER_PER_TOKEN * tokens,
                "note": "algorithmic FLOPs; every product is a 3-term split-fp16 MMA (parity), so frac <= 1/3",
                "launch_ms": layer_ms, "launches_timed": n_layer,
                "share_of_step": layer_ms * 8 / (serial_ms / K), "serial_ms_per_step": serial_ms / K,
                "measured": "profiler pass after the timed region: one stream, whole batch per launch (the timed region "
                            "itself interleaves sub-batches on several streams)",
                "whole_path_tflops": FLOPS_PER_PAIR * B * K / (ms * 1e-3) / 1e12}
    else:
        ach = FLOPS_PER_PAIR * B * K / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "whole hot path (fp32 CUDA-core path has no single dominant kernel)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src}
    cpu_pps, cpu_dt = cpu_port_pairs_per_s(2, 4)
    # the same path as eager PyTorch fp32 (TF32 off) on this GPU: the "GPU bar to beat" of SURVEY.md 8(d)
    eager = None
    try:
        from oracle import oetr_torch_eager as ote
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        Wt = ote.prepare(Wts, dev)
        for _ in range(3):
            t1, t2 = ote.hot_path(Wt, feats[0][0], feats[0][1], (IMG, IMG), (IMG, IMG))
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev4.record()
        for i in range(5):
            t1, t2 = ote.hot_path(Wt, feats[i % N_ROTATE][0], feats[i % N_ROTATE][1], (IMG, IMG), (IMG, IMG))
        ev5.record()
        torch.cuda.synchronize()
        g1, g2 = hot.forward(feats[4][0], feats[4][1], (IMG, IMG), (IMG, IMG))
        torch.cuda.synchronize()
        eager = {"value": B * 5 / (ev4.elapsed_time(ev5) * 1e-3), "unit": "pairs/s",
                 "kind": "eager PyTorch fp32 (TF32 off) restatement of the same path (oracle/oetr_torch_eager.py), "
                         "same GPU, same batch, device-resident inputs, 5 steps",
                 "max_box_diff_px": float(max((g1 - t1).abs().max().item(), (g2 - t2).abs().max().item()))}
        del Wt
    except Exception as e:  # measurement aid only
        eager = {"error": repr(e)}
    value = world * B * K / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": "batch=32 640x640 pairs per GPU: hot path (8-layer correlation transformer + "
                               "decoder + overlap head) on two [32,256,20,20] fp32 feature maps",
                   "pairs_per_gpu": B, "precision": args.precision, "attention": "linear",
                   "sub_batches": ("%d pairs each, own stream" % (args.chunk_pairs if args.chunk_pairs > 0 else 8)) if args.chunk_pairs else "off",
                   "batches_in_flight": "%d (one CUDA stream + handle each; every step = one complete forward)" % lanes,
                   "l2": "inputs rotate over %d resident batches (%.0f MB > 126 MB L2)" % (
                       N_ROTATE, N_ROTATE * 2 * base1.nbytes / 1e6),
                   "parallelism": "batch shards, replicated weights, all-gather of boxes" if world > 1 else "1 GPU"},
        "clocks": clk.summary(),
        "e2e": {"value": world * B * K / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * base1.nbytes),
                "d2h_bytes_per_step": int(B * 8 * 4), "api": "oetr_forward_host_submit/_wait (C ABI, pinned host feature buffers, %d requests in flight)" % depth,
                "blocking_value": world * B * K / e2e_blocking_s,
                "blocking_api": "oetr_forward_host (one request at a time; rank-local time)"},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": {"value": cpu_pps, "unit": "pairs/s", "cores": _host_threads(), "kind": "port",
                         "sample": "2 pairs x 4 runs of the numpy fp32 port (%.1f s)" % cpu_dt},
        "gpu_eager_baseline": eager,
        "parity": {"box_err_over_image_side": float(perr), "bar": 1e-3, "pairs_checked": 2},
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunk-pairs", type=int, default=int(os.environ.get("OETR_CHUNK_PAIRS", "-1")),
                    help="pairs per concurrently scheduled sub-batch (0 = off, -1 = automatic: 8 at 640x640)")
    ap.add_argument("--in-flight", type=int, default=int(os.environ.get("OETR_IN_FLIGHT", "2")),
                    help="independent batches kept in flight in the device-resident timed region")
    ap.add_argument("--e2e-in-flight", type=int, default=4, help="host requests kept in flight in the e2e leg (1..4)")
    ap.add_argument("--precision", default=os.environ.get("OETR_BENCH_PRECISION", "fp16"), choices=["fp16", "fp32"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
