"""Per-SM busy time of the big kernels under the production schedule, from the CTA log of the library (OETR_TIMING=2):
every k_enc / k_conv CTA appends {smid, kernel, t_start, t_end} (globaltimer ns).  Prints, for the timed window, the
fraction of SM-time covered by each kernel class and the idle fraction.  Usage: OETR_TIMING=2 python tools/sm_timeline.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oetr_b200  # noqa: E402
from oetr_b200 import cabi, weights  # noqa: E402

NAMES = {1: "k_enc source phase (first launch)", 2: "k_enc layer (query + source phase)", 3: "k_enc decoder K/V", 4: "k_conv"}


def main():
    os.environ["OETR_TIMING"] = "2"
    inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    W = weights.synthetic_hot_path_weights(0)
    hots = [oetr_b200.OverlapHotPath(W) for _ in range(inflight)]
    streams = [torch.cuda.Stream() for _ in hots]
    f1 = torch.from_numpy(weights.synthetic_features(32, 20, 20, seed=1, tag="a")).cuda()
    f2 = torch.from_numpy(weights.synthetic_features(32, 20, 20, seed=1, tag="b")).cuda()
    lib = cabi.load_library()

    def run(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % inflight]):
                hots[i % inflight].forward(f1, f2, (640, 640), (640, 640))
        torch.cuda.synchronize()

    run(6)
    cap = 48 + 8 + (1 << 18) * 4
    buf = (ctypes.c_ulonglong * cap)()
    lib.oetr_debug_cycles(buf, 48, 1)
    n0 = buf[47]
    steps = 30
    run(steps)
    lib.oetr_debug_cycles(buf, cap, 0)
    v = np.frombuffer(buf, dtype=np.uint64)
    n = int(v[47])
    log = v[56:56 + 4 * min(n, 1 << 18)].reshape(-1, 4).astype(np.int64)
    log = log[int(n0):]                                   # entries of the timed run only
    t0, t1 = log[:, 2].min(), log[:, 3].max()
    # trim the ramp-up / ramp-down: the middle 80 % of the window
    a, b = t0 + (t1 - t0) // 10, t1 - (t1 - t0) // 10
    sms = np.unique(log[:, 0])
    print("%d CTAs on %d SMs, window %.2f ms (middle 80 %% of %d steps)" % (len(log), len(sms), (b - a) * 1e-6, steps))
    total = (b - a) * len(sms)
    busy_all = 0
    for k, name in NAMES.items():
        e = log[log[:, 1] == k]
        busy = np.clip(np.minimum(e[:, 3], b) - np.maximum(e[:, 2], a), 0, None).sum()
        busy_all += busy
        dur = (e[:, 3] - e[:, 2]).mean() * 1e-3 if len(e) else 0.0
        print("  %-40s %6d CTAs, mean %.1f us, %5.1f %% of SM-time" % (name, len(e), dur, 100.0 * busy / total))
    print("  %-40s %5.1f %% (small kernels, launch gaps, dependency stalls)" % ("not covered by these kernels", 100.0 * (1 - busy_all / total)))


if __name__ == "__main__":
    main()
