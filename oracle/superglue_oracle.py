"""CPU oracle of SuperGlue -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy (fp64) restatement of third_party/SuperGluePretrainedNetwork/models/superglue.py of the reference tree: attention
(:86-90), MultiHeadedAttention (:93-108), AttentionalPropagation (:111-120), AttentionalGNN (:123-140),
log_sinkhorn_iterations / log_optimal_transport (:143-184) and SuperGlue.forward (:236-290), batch of one.
Weights: {state_dict key -> ndarray} with the reference's names (superglue_{indoor,outdoor}.pth load as they are).
Parity status: pinned against the UNMODIFIED reference module with its in-tree outdoor weights
(tests/golden/make_superglue_golden.py -> tests/golden/superglue_*.npz; tests/test_superglue_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
"""
import numpy as np

BN_EPS = 1e-5


def conv1d(W, name, x):
    """nn.Conv1d(kernel_size=1): x [C_in, N] -> [C_out, N]"""
    return W[name + ".weight"][:, :, 0] @ x + W[name + ".bias"][:, None]


def batch_norm(W, name, x):
    """nn.BatchNorm1d in eval mode"""
    g, b, mu, var = (W[name + s][:, None] for s in (".weight", ".bias", ".running_mean", ".running_var"))
    return (x - mu) / np.sqrt(var + BN_EPS) * g + b


def mlp(W, prefix, n_layers, x):
    """MLP(channels) (:51-62): Conv1d, then BatchNorm1d + ReLU after every layer but the last; Sequential indices 0,1,2 | 3,.."""
    for i in range(n_layers):
        x = conv1d(W, "%s.%d" % (prefix, 3 * i), x)
        if i < n_layers - 1:
            x = np.maximum(batch_norm(W, "%s.%d" % (prefix, 3 * i + 1), x), 0.0)
    return x


def normalize_keypoints(kpts, image_shape):
    """:65-72; kpts [N,2] (x, y), image_shape (.., height, width)"""
    height, width = image_shape[-2:]
    size = np.asarray([width, height], dtype=np.float64)
    return (kpts - size / 2) / (size.max() * 0.7)


def attention(query, key, value):
    """:86-90 for one batch element: query [dim, heads, n], key / value [dim, heads, m] -> [dim, heads, n]"""
    dim = query.shape[0]
    scores = np.einsum("dhn,dhm->hnm", query, key) / dim ** 0.5
    scores = scores - scores.max(axis=-1, keepdims=True)
    prob = np.exp(scores)
    prob /= prob.sum(axis=-1, keepdims=True)
    return np.einsum("hnm,dhm->dhn", prob, value)


def multi_head_attention(W, prefix, query, key, value, heads=4):
    """:100-108; the view(batch, dim, heads, -1) makes channel c = d * heads + h"""
    d_model = query.shape[0]
    q, k, v = (conv1d(W, "%s.proj.%d" % (prefix, i), x).reshape(d_model // heads, heads, -1)
               for i, x in enumerate((query, key, value)))
    return conv1d(W, prefix + ".merge", attention(q, k, v).reshape(d_model, -1))


def attentional_propagation(W, prefix, x, source):
    message = multi_head_attention(W, prefix + ".attn", x, source, source)
    return mlp(W, prefix + ".mlp", 2, np.concatenate([x, message], axis=0))


def attentional_gnn(W, desc0, desc1, names=("self", "cross") * 9):
    for i, name in enumerate(names):
        src0, src1 = (desc1, desc0) if name == "cross" else (desc0, desc1)
        p = "gnn.layers.%d" % i
        d0, d1 = attentional_propagation(W, p, desc0, src0), attentional_propagation(W, p, desc1, src1)
        desc0, desc1 = desc0 + d0, desc1 + d1
    return desc0, desc1


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def log_optimal_transport(scores, alpha, iters):
    """:150-184 for one batch element: scores [m, n] -> [m+1, n+1]"""
    m, n = scores.shape
    Z = np.full((m + 1, n + 1), float(alpha), dtype=scores.dtype)
    Z[:m, :n] = scores
    norm = -np.log(m + n)
    log_mu = np.concatenate([np.full(m, norm), [np.log(n) + norm]])
    log_nu = np.concatenate([np.full(n, norm), [np.log(m) + norm]])
    u, v = np.zeros(m + 1), np.zeros(n + 1)
    for _ in range(iters):
        u = log_mu - _lse(Z + v[None, :], 1)
        v = log_nu - _lse(Z + u[:, None], 0)
    return Z + u[:, None] + v[None, :] - norm


def match(scores, threshold):
    """:268-282 -> (matches0 [m], matches1 [n], mscores0, mscores1)"""
    inner = scores[:-1, :-1]
    i0, i1 = inner.argmax(1), inner.argmax(0)
    mutual0 = np.arange(len(i0)) == i1[i0]
    mutual1 = np.arange(len(i1)) == i0[i1]
    ms0 = np.where(mutual0, np.exp(inner.max(1)), 0.0)
    ms1 = np.where(mutual1, ms0[i1], 0.0)
    valid0 = mutual0 & (ms0 > threshold)
    valid1 = mutual1 & valid0[i1]
    return np.where(valid0, i0, -1), np.where(valid1, i1, -1), ms0, ms1


def superglue(W, kpts0, kpts1, scores0, scores1, desc0, desc1, shape0, shape1, iters=100, threshold=0.2, dtype=np.float64):
    """SuperGlue.forward (:236-290), batch of one: kpts [N,2], scores [N], desc [256,N], image shapes (.., H, W)."""
    W = {k: np.asarray(v, dtype=dtype) for k, v in W.items() if not k.endswith("num_batches_tracked")}
    k0, k1 = normalize_keypoints(np.asarray(kpts0, dtype), shape0), normalize_keypoints(np.asarray(kpts1, dtype), shape1)
    enc = lambda k, s: mlp(W, "kenc.encoder", 5, np.concatenate([k.T, np.asarray(s, dtype)[None]], axis=0))
    d0, d1 = np.asarray(desc0, dtype) + enc(k0, scores0), np.asarray(desc1, dtype) + enc(k1, scores1)
    d0, d1 = attentional_gnn(W, d0, d1)
    m0, m1 = conv1d(W, "final_proj", d0), conv1d(W, "final_proj", d1)
    sc = m0.T @ m1 / 256 ** 0.5
    Z = log_optimal_transport(sc, float(W["bin_score"]), iters)
    out = match(Z, threshold)
    return {"scores": Z, "matches0": out[0], "matches1": out[1], "matching_scores0": out[2], "matching_scores1": out[3],
            "pre_transport": sc, "desc0": d0, "desc1": d1}
