"""SuperGlue with its two hot operators on the device: same constructor config, parameter names (the
reference's superglue_{indoor,outdoor}.pth load strictly) and forward dict as the reference's
third_party/SuperGluePretrainedNetwork/models/superglue.py:186-290.  The 1x1 Conv1d / BatchNorm layers stay PyTorch (cuBLAS:
plain library GEMMs); `attention` (:86-90) and `log_optimal_transport` (:150-184) are CUDA kernels of liboetr_b200.so
(oetr_sg_attention: online softmax, no [N, M] matrix in memory; oetr_sg_optimal_transport: dustbin-augmented log-Sinkhorn
without materialising the augmented matrix).  Inference only; no CPU or PyTorch fallback for the two operators."""
import ctypes
from copy import deepcopy

import torch
from torch import nn

from . import cabi


def _check(rc, lib):
    if rc != cabi.OETR_OK:
        raise cabi.OetrError(rc, (lib.oetr_sg_last_error() or b"").decode("utf-8", "replace"))


def _need_cuda(t, what):
    if not t.is_cuda:
        raise cabi.OetrError(cabi.OETR_E_ARCH, "%s on %s: the SuperGlue operators have no CPU fallback" % (what, t.device))


MODES = {"tensor": 0, "fp32": 1}


def attention(query, key, value, mode="tensor"):
    """reference superglue.py:86-90: query [b, dim=64, heads=4, n], key / value [b, 64, 4, m] -> [b, 64, 4, n].  The
    probability tensor the reference also returns is never used by its callers and is not produced.  mode: "tensor" =
    tcgen05 with split fp16 operands (default), "fp32" = the CUDA-core kernel."""
    _need_cuda(query, "attention")
    b, d, h, n = query.shape
    m = key.shape[3]
    if (d, h) != (64, 4) or key.shape[:3] != (b, 64, 4) or value.shape != key.shape:
        raise ValueError("attention is specialised for 4 heads x 64 dims, got %s / %s / %s" % (tuple(query.shape), tuple(key.shape), tuple(value.shape)))
    q, k, v = (t.contiguous().float() for t in (query, key, value))
    out = torch.empty_like(q)
    lib = cabi.load_library()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(q.device):
        _check(lib.oetr_sg_attention(vp(q), vp(k), vp(v), vp(out), b, n, m, MODES[mode], ctypes.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)), lib)
    return out


def log_optimal_transport(scores, alpha, iters):
    """reference superglue.py:150-184: scores [b, m, n], alpha = bin_score -> [b, m+1, n+1]"""
    _need_cuda(scores, "log_optimal_transport")
    b, m, n = scores.shape
    s = scores.contiguous().float()
    out = torch.empty(b, m + 1, n + 1, dtype=torch.float32, device=s.device)
    lib = cabi.load_library()
    ws = torch.empty(lib.oetr_sg_transport_workspace_bytes(b, m, n), dtype=torch.uint8, device=s.device)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(s.device):
        _check(lib.oetr_sg_optimal_transport(vp(s), float(alpha), int(iters), vp(out), b, m, n, vp(ws), ws.numel(),
                                             ctypes.c_void_p(torch.cuda.current_stream(s.device).cuda_stream)), lib)
    return out


def MLP(channels, do_bn=True):
    layers = []
    for i in range(1, len(channels)):
        layers.append(nn.Conv1d(channels[i - 1], channels[i], kernel_size=1, bias=True))
        if i < len(channels) - 1:
            if do_bn:
                layers.append(nn.BatchNorm1d(channels[i]))
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


def normalize_keypoints(kpts, image_shape):
    _, _, height, width = image_shape
    size = kpts.new_tensor([[float(width), float(height)]])
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center[:, None, :]) / scaling[:, None, :]


class KeypointEncoder(nn.Module):
    def __init__(self, feature_dim, layers):
        super().__init__()
        self.encoder = MLP([3] + list(layers) + [feature_dim])

    def forward(self, kpts, scores):
        return self.encoder(torch.cat([kpts.transpose(1, 2), scores.unsqueeze(1)], dim=1))


class MultiHeadedAttention(nn.Module):
    def __init__(self, num_heads, d_model):
        super().__init__()
        self.dim, self.num_heads = d_model // num_heads, num_heads
        self.merge = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.proj = nn.ModuleList([deepcopy(self.merge) for _ in range(3)])

    def forward(self, query, key, value):
        b = query.size(0)
        query, key, value = [l(x).view(b, self.dim, self.num_heads, -1) for l, x in zip(self.proj, (query, key, value))]
        x = attention(query, key, value)
        return self.merge(x.view(b, self.dim * self.num_heads, -1))


class AttentionalPropagation(nn.Module):
    def __init__(self, feature_dim, num_heads):
        super().__init__()
        self.attn = MultiHeadedAttention(num_heads, feature_dim)
        self.mlp = MLP([feature_dim * 2, feature_dim * 2, feature_dim])

    def forward(self, x, source):
        return self.mlp(torch.cat([x, self.attn(x, source, source)], dim=1))


class AttentionalGNN(nn.Module):
    """reference superglue.py:123-140.  Every layer's Conv1d / BatchNorm1d is pointwise along the keypoint axis and both
    images use the same weights, so the two descriptor sets travel as ONE tensor [b, 256, n0 + n1]: per layer three
    projections, one merge and one MLP call instead of twice as many (the forward is launch-bound: ~10 small kernels per
    layer instead of ~24), and only the attention itself runs per image (self: own keys; cross: the other image's keys)."""

    def __init__(self, feature_dim, layer_names):
        super().__init__()
        self.layers = nn.ModuleList([AttentionalPropagation(feature_dim, 4) for _ in layer_names])
        self.names = list(layer_names)

    def forward(self, desc0, desc1):
        b, n0 = desc0.size(0), desc0.size(2)
        x = torch.cat([desc0, desc1], dim=2)
        for layer, name in zip(self.layers, self.names):
            att = layer.attn
            q, k, v = [proj(x).view(b, att.dim, att.num_heads, -1) for proj in att.proj]
            q0, q1 = q[..., :n0].contiguous(), q[..., n0:].contiguous()
            k0, k1 = k[..., :n0].contiguous(), k[..., n0:].contiguous()
            v0, v1 = v[..., :n0].contiguous(), v[..., n0:].contiguous()
            if name == "cross":
                m0, m1 = attention(q0, k1, v1), attention(q1, k0, v0)
            else:
                m0, m1 = attention(q0, k0, v0), attention(q1, k1, v1)
            message = att.merge(torch.cat([m0, m1], dim=3).view(b, att.dim * att.num_heads, -1))
            x = x + layer.mlp(torch.cat([x, message], dim=1))
        return x[..., :n0], x[..., n0:]


class SuperGlue(nn.Module):
    default_config = {
        "descriptor_dim": 256,
        "weights": None,                  # a path to superglue_{indoor,outdoor}.pth, or None (load_state_dict yourself)
        "keypoint_encoder": [32, 64, 128, 256],
        "GNN_layers": ["self", "cross"] * 9,
        "sinkhorn_iterations": 100,
        "match_threshold": 0.2,
    }

    def __init__(self, config=None):
        super().__init__()
        self.config = {**self.default_config, **(config or {})}
        if self.config["descriptor_dim"] != 256:
            raise ValueError("the CUDA attention is specialised for descriptor_dim 256 (4 heads x 64)")
        self.kenc = KeypointEncoder(256, self.config["keypoint_encoder"])
        self.gnn = AttentionalGNN(256, self.config["GNN_layers"])
        self.final_proj = nn.Conv1d(256, 256, kernel_size=1, bias=True)
        self.register_parameter("bin_score", nn.Parameter(torch.tensor(1.0)))
        if self.config["weights"]:
            self.load_state_dict(torch.load(str(self.config["weights"]), map_location="cpu"))

    @torch.no_grad()
    def forward(self, data):
        desc0, desc1 = data["descriptors0"], data["descriptors1"]
        kpts0, kpts1 = data["keypoints0"], data["keypoints1"]
        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:
            shape0, shape1 = kpts0.shape[:-1], kpts1.shape[:-1]
            return {"matches0": kpts0.new_full(shape0, -1, dtype=torch.int), "matches1": kpts1.new_full(shape1, -1, dtype=torch.int),
                    "matching_scores0": kpts0.new_zeros(shape0), "matching_scores1": kpts1.new_zeros(shape1)}
        kpts0 = normalize_keypoints(kpts0, data["image0"].shape)
        kpts1 = normalize_keypoints(kpts1, data["image1"].shape)
        desc0 = desc0 + self.kenc(kpts0, data["scores0"])
        desc1 = desc1 + self.kenc(kpts1, data["scores1"])
        desc0, desc1 = self.gnn(desc0, desc1)
        mdesc0, mdesc1 = self.final_proj(desc0), self.final_proj(desc1)
        scores = torch.einsum("bdn,bdm->bnm", mdesc0, mdesc1) / 256 ** 0.5
        scores = log_optimal_transport(scores, self.bin_score, iters=self.config["sinkhorn_iterations"])
        m0, m1, s0, s1 = mutual_matches(scores, self.config["match_threshold"])
        return {"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1, "scores": scores}


def mutual_matches(log_assignment, threshold):
    """The decision rule of reference superglue.py:268-282 on the [b, m+1, n+1] log assignment matrix: keypoint i of image
    0 and j of image 1 match when each is the other's arg-max over the non-dustbin block and exp(score) exceeds the
    threshold; -1 marks unmatched keypoints.  Returns (matches0 [b,m], matches1 [b,n], scores0, scores1)."""
    block = log_assignment[:, :-1, :-1]
    best_for_0 = block.max(dim=2)                     # per keypoint of image 0: best partner in image 1
    best_for_1 = block.max(dim=1)
    j_of_i, i_of_j = best_for_0.indices, best_for_1.indices
    rows = torch.arange(j_of_i.shape[1], device=block.device).expand_as(j_of_i)
    cols = torch.arange(i_of_j.shape[1], device=block.device).expand_as(i_of_j)
    agree0 = i_of_j.gather(1, j_of_i) == rows          # i -> j -> back to i
    agree1 = j_of_i.gather(1, i_of_j) == cols
    conf0 = best_for_0.values.exp() * agree0
    conf1 = conf0.gather(1, i_of_j) * agree1
    keep0 = agree0 & (conf0 > threshold)
    keep1 = agree1 & keep0.gather(1, i_of_j)
    unmatched = j_of_i.new_full((), -1)
    return torch.where(keep0, j_of_i, unmatched), torch.where(keep1, i_of_j, unmatched), conf0, conf1
