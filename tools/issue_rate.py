"""Is the device-resident timed region of bench.py bound by the host's launch rate?  Times how long the host needs to ISSUE
K forwards (no synchronisation) against how long the GPU needs to run them.   python tools/issue_rate.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oetr_b200  # noqa: E402
from oetr_b200 import weights  # noqa: E402

W = weights.synthetic_hot_path_weights(0)
lanes = 3
hots = [oetr_b200.OverlapHotPath(W, precision="fp16") for _ in range(lanes)]
streams = [torch.cuda.Stream() for _ in range(lanes)]
f1 = torch.from_numpy(weights.synthetic_features(32, 20, 20, seed=1, tag="a")).cuda()
f2 = torch.from_numpy(weights.synthetic_features(32, 20, 20, seed=1, tag="b")).cuda()


def run(K):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        with torch.cuda.stream(streams[i % lanes]):
            hots[i % lanes].forward(f1, f2, (640, 640), (640, 640))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3


run(6)
for K in (20, 60):
    issue, total = run(K)
    print("K=%d: host issue %.3f ms/step, total %.3f ms/step (launches per forward %d)" % (K, issue, total, hots[0].last_launch_count))
