"""Where a k_enc tile spends its cycles UNDER THE PRODUCTION SCHEDULE (sub-batches on several streams, two batches in
flight): run with OETR_TIMING=1.  The kernels accumulate cycle counts with atomicAdd (no host synchronisation):
   MMA lane: total / waiting for the operand image / waiting for weights     (per tile of a q+kv launch)
   row warp thread 0: duration of every stage between two stamps (tc_enc.cuh `stamp(i)`)
Usage: OETR_TIMING=1 python tools/stage_cycles.py [--batch 32] [--side 640] [--steps 20] [--in-flight 2]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oetr_b200  # noqa: E402
from oetr_b200 import cabi, weights  # noqa: E402

STAGES = ["x load", "E0 LNq image", "wait q", "E1 phi(q)/Z", "wait msg", "E2 x+=msg, LN2 image", "wait h_a",
          "E3 gelu(h_a) (+wait h_b)", "E4 gelu(h_b) (+wait y_a)", "wait y", "E5 x+=y", "store x", "LNkv image",
          "wait v,k", "kv epilogue (+wait KV half 0)", "wait KV", "KV out"]
NS = 48


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--side", type=int, default=640)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--in-flight", type=int, default=2)
    ap.add_argument("--chunk-pairs", type=int, default=-1)
    a = ap.parse_args()
    fm = a.side // 32
    W = weights.synthetic_hot_path_weights(0)
    hots = [oetr_b200.OverlapHotPath(W) for _ in range(a.in_flight)]
    for h in hots:
        h.set_chunk_pairs(a.chunk_pairs)
    streams = [torch.cuda.Stream() for _ in hots]
    f1 = torch.from_numpy(weights.synthetic_features(a.batch, fm, fm, seed=1, tag="a")).cuda()
    f2 = torch.from_numpy(weights.synthetic_features(a.batch, fm, fm, seed=1, tag="b")).cuda()
    lib = cabi.load_library()
    buf = (ctypes.c_ulonglong * NS)()

    def run(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % len(hots)]):
                hots[i % len(hots)].forward(f1, f2, (a.side, a.side), (a.side, a.side))
        torch.cuda.synchronize()

    run(4)
    lib.oetr_debug_cycles(buf, NS, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run(a.steps)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.steps
    n = lib.oetr_debug_cycles(buf, NS, 1)
    v = np.array(list(buf), dtype=np.float64)
    print("batch %d, %dx%d, %d in flight: %.3f ms/step, %.0f pairs/s (timing build perturbs little: atomics only)" % (
        a.batch, a.side, a.side, a.in_flight, ms, a.batch / ms * 1e3))
    if n == 0 or v[3] == 0:
        print("no timing data (OETR_TIMING not set when the library was first used?)")
        return
    t = v[3]
    print("k_enc q+kv launches: %d tiles sampled; per tile (cycles):" % t)
    print("  MMA lane total %8.0f   waiting operand image %8.0f   waiting weights %8.0f   issuing/executing %8.0f" % (
        v[0] / t, v[1] / t, v[2] / t, (v[0] - v[1] - v[2]) / t))
    tot = v[8:8 + len(STAGES)].sum() / t
    print("  row warp 0 stages (sum %.0f):" % tot)
    for i, name in enumerate(STAGES):
        print("    %2d %-34s %8.0f  %5.1f %%" % (i, name, v[8 + i] / t, 100 * v[8 + i] / t / tot))
    if v[43]:
        c = v[43]
        print("k_conv: %d tiles; MMA lane total %.0f, waiting operand image %.0f, waiting weights %.0f" % (
            c, v[40] / c, v[41] / c, v[42] / c))


if __name__ == "__main__":
    main()
