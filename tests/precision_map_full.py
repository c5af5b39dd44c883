"""Operand-precision experiment for the FULL-attention tensor-core path (attention_mode='full', reference
src/models/linear_attention.py:59-87) -- TEST INFRASTRUCTURE, CPU only.  Same method as tests/precision_map.py: every GEMM
operand is either exact ("x": hi+lo split) or ONE fp16 value ("h"); the box error of each single choice and of candidate
maps is measured on the full-attention golden geometries and stress scales.  Run: python tests/precision_map_full.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oetr_oracle as orc  # noqa: E402
from oetr_b200 import weights  # noqa: E402
import precision_map as pm  # noqa: E402

SITES = ["a_q", "w_q", "a_kv", "w_k", "w_v", "q_op", "k_op", "p_op", "v_op", "a_m", "w_m", "a_1", "w_1", "a_2", "w_2"]


def enc_layer(W, pre, x, s, xp, sp, m):
    g = lambda n: W[pre + n]
    n, L, _ = x.shape
    S = s.shape[1]
    qi = orc.layer_norm(x, g("pre_norm_q.weight"), g("pre_norm_q.bias")) + xp
    kvi = orc.layer_norm(s, g("pre_norm_kv.weight"), g("pre_norm_kv.bias")) + sp
    q = (m.r("a_q", qi) @ m.r("w_q", g("q_proj.weight")).T).reshape(n, L, 8, 32)
    k = (m.r("a_kv", kvi) @ m.r("w_k", g("k_proj.weight")).T).reshape(n, S, 8, 32)
    v = (m.r("a_kv", kvi) @ m.r("w_v", g("v_proj.weight")).T).reshape(n, S, 8, 32)
    # the kernel folds 1/sqrt(32) * log2(e) into the Q operand; rounding is scale-free except for that factor's own rounding
    qs = m.r("q_op", q / 32 ** 0.5)
    A = np.einsum("nlhd,nshd->nhls", qs, m.r("k_op", k))
    A = A - A.max(axis=3, keepdims=True)
    P = np.exp(A)
    l = P.sum(axis=3, keepdims=True)                       # the row sum uses the unrounded P (fp32 in the kernel)
    O = np.einsum("nhls,nshd->nlhd", m.r("p_op", P), m.r("v_op", v)) / l.transpose(0, 2, 1, 3)
    msg = m.r("a_m", O.reshape(n, L, 256)) @ m.r("w_m", g("merge.weight")).T
    x = x + msg
    h = orc.gelu_erf(m.r("a_1", orc.layer_norm(x, g("norm2.weight"), g("norm2.bias"))) @ m.r("w_1", g("mlp.0.weight")).T)
    return x + m.r("a_2", h) @ m.r("w_2", g("mlp.2.weight")).T


def main():
    pm.enc_layer = enc_layer                                # the driver of precision_map with the full-attention layer
    cases = [c for c in pm.make_cases() if c[0] in ("b2_640", "ragged", "tiny_b3", "feat_x3", "ln_gain3")]
    refs = [pm.run(W, f1, f2, hw1, hw2, pm.Map()) for (_, W, f1, f2, hw1, hw2) in cases]
    names = [c[0] for c in cases]
    print("%-8s" % "site", " ".join("%9s" % n for n in names), "      max")
    for s in SITES:
        es = [pm.err(pm.run(W, f1, f2, hw1, hw2, pm.Map({s: "h"})), r) for (_, W, f1, f2, hw1, hw2), r in zip(cases, refs)]
        print("%-8s" % s, " ".join("%9.2e" % e for e in es), "%9.2e" % max(es), flush=True)
    for label, hs in CANDIDATES.items():
        mp = pm.Map({s: "h" for s in hs})
        es = [pm.err(pm.run(W, f1, f2, hw1, hw2, mp), r) for (_, W, f1, f2, hw1, hw2), r in zip(cases, refs)]
        print("%-34s" % label, " ".join("%9.2e" % e for e in es), "%9.2e" % max(es), flush=True)


CANDIDATES = {
    "attention operands h (q,k,p,v)": ["q_op", "k_op", "p_op", "v_op"],
    "q,k,p h; v split": ["q_op", "k_op", "p_op"],
    "q_op,k_op h": ["q_op", "k_op"],
    "full map A: a_q,w_q,w_k,q,k,p,v h": ["a_q", "w_q", "w_k", "q_op", "k_op", "p_op", "v_op"],
}

if __name__ == "__main__":
    main()
