"""Debug driver: one forward of the fp16 path at a given batch / chunking, compared with the unsplit k_enc result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oetr_b200
from oetr_b200 import weights

b = int(sys.argv[1]); cp = int(sys.argv[2]); fm = int(sys.argv[3]) if len(sys.argv) > 3 else 20
W = weights.synthetic_hot_path_weights(0)
f1 = torch.from_numpy(weights.synthetic_features(b, fm, fm, seed=9, tag="p1")).cuda()
f2 = torch.from_numpy(weights.synthetic_features(b, fm, fm, seed=9, tag="p2")).cuda()
hot = oetr_b200.OverlapHotPath(W, precision="fp16")
hot.set_chunk_pairs(cp)
t0 = time.time()
a1, a2 = hot.forward(f1, f2, (fm * 32, fm * 32), (fm * 32, fm * 32), clamp=False)
torch.cuda.synchronize()
print("forward done in %.3f s" % (time.time() - t0), flush=True)
try:
    hot.poll_error()
    print("poll_error: ok", flush=True)
except Exception as e:
    print("poll_error:", e, flush=True)
ref = oetr_b200.OverlapHotPath(W, precision="fp32")
r1, r2 = ref.forward(f1, f2, (fm * 32, fm * 32), (fm * 32, fm * 32), clamp=False)
torch.cuda.synchronize()
print("max box err / side vs fp32 path: %.3e %.3e" % ((a1 - r1).abs().max().item() / (fm * 32), (a2 - r2).abs().max().item() / (fm * 32)), flush=True)
for i in range(3):
    t0 = time.time()
    hot.forward(f1, f2, (fm * 32, fm * 32), (fm * 32, fm * 32), clamp=False)
    torch.cuda.synchronize()
    print("repeat %d: %.3f ms" % (i, (time.time() - t0) * 1e3), flush=True)
