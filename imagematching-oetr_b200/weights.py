"""Hot-path weight inventory, canonical packing for the C ABI, and deterministic synthetic weights.

The names are the reference's state_dict keys (reference src/model.py:58-86,
src/models/transformer.py:23-311).  `CANONICAL_ORDER` is the order documented in include/oetr_b200.h.
"""
import zlib

import numpy as np

D_MODEL = 256


def _encoder_names(i):
    p = "transformer.encoder.%d." % i
    return [
        (p + "q_proj.weight", (256, 256)), (p + "k_proj.weight", (256, 256)), (p + "v_proj.weight", (256, 256)),
        (p + "merge.weight", (256, 256)), (p + "mlp.0.weight", (512, 256)), (p + "mlp.2.weight", (256, 512)),
        (p + "pre_norm_q.weight", (256,)), (p + "pre_norm_q.bias", (256,)),
        (p + "pre_norm_kv.weight", (256,)), (p + "pre_norm_kv.bias", (256,)),
        (p + "norm2.weight", (256,)), (p + "norm2.bias", (256,)),
    ]


def _decoder_names(j):
    p = "transformer.decoder.layers.%d." % j
    out = []
    for a in ("self_attn.", "multihead_attn."):
        for proj in ("q_proj", "k_proj", "v_proj"):
            out += [(p + a + proj + ".weight", (256, 256)), (p + a + proj + ".bias", (256,))]
        out.append((p + a + "merge.weight", (256, 256)))
    out += [(p + "mlp.0.weight", (512, 256)), (p + "mlp.2.weight", (256, 512))]
    for n in ("norm1", "norm2", "norm3"):
        out += [(p + n + ".weight", (256,)), (p + n + ".bias", (256,))]
    return out


def _unused_decoder_names(j):
    # present in every reference checkpoint, never read by forward (transformer.py:197-202)
    p = "transformer.decoder.layers.%d." % j
    return [(p + n + ".weight", (256, 256)) for n in ("q_proj", "k_proj", "v_proj", "merge")]


CANONICAL_ORDER = (
    [t for i in range(8) for t in _encoder_names(i)]
    + [t for j in range(2) for t in _decoder_names(j)]
    + [("query_embed1.weight", (1, 256)), ("query_embed2.weight", (1, 256)),
       ("tlbr_reg.0.weight", (256, 256)), ("tlbr_reg.2.weight", (4, 256)), ("tlbr_reg.2.bias", (4,)),
       ("heatmap_conv.0.weight", (256, 256, 3, 3)), ("heatmap_conv.0.bias", (256,)),
       ("heatmap_conv.1.weight", (256,)), ("heatmap_conv.1.bias", (256,)),
       ("heatmap_conv.3.weight", (1, 256, 1, 1)), ("heatmap_conv.3.bias", (1,))]
)
UNUSED_NAMES = [t for j in range(2) for t in _unused_decoder_names(j)]
PACKED_COUNT = sum(int(np.prod(s)) for _, s in CANONICAL_ORDER)     # 6 443 525


def pack_hot_path_weights(state_dict):
    """state_dict (torch tensors or ndarrays, reference key names) -> one fp32 ndarray in canonical order."""
    parts = []
    for name, shape in CANONICAL_ORDER:
        if name not in state_dict:
            raise KeyError("hot-path weight %r missing from state dict" % name)
        v = state_dict[name]
        if hasattr(v, "detach"):
            v = v.detach().to("cpu").float().numpy()
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(shape):
            raise ValueError("weight %r has shape %s, expected %s" % (name, tuple(v.shape), shape))
        parts.append(v.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def _rng_for(name, seed):
    return np.random.Generator(np.random.Philox(key=[zlib.crc32(name.encode()) & 0xFFFFFFFF, seed]))


def _uniform(rng, shape, bound):
    # 24-bit uniform built from raw integers: reproducible across numpy versions
    u = rng.integers(0, 1 << 24, size=shape, dtype=np.int64).astype(np.float64) / float(1 << 24)
    return ((2.0 * u - 1.0) * bound).astype(np.float32)


def synthetic_tensor(name, shape, seed=0):
    """Deterministic stand-in for a trained tensor, keyed by its state_dict name.
    Matrices: xavier-uniform bounds like transformer.py:308-311; LayerNorm/GroupNorm scales around 1 and all
    biases non-zero so that every affine term is exercised (default inits of 1/0 would hide bugs)."""
    rng = _rng_for(name, seed)
    shape = tuple(shape)
    if name == "tlbr_reg.2.weight":
        return _uniform(rng, shape, 0.02)          # keeps sigmoid(tlbr) away from saturation: a sharper test
    if len(shape) >= 2 and int(np.prod(shape)) > 512:
        fan_out = shape[0] * int(np.prod(shape[2:]))
        fan_in = shape[1] * int(np.prod(shape[2:]))
        return _uniform(rng, shape, float(np.sqrt(6.0 / (fan_in + fan_out))))
    if "query_embed" in name:
        return _uniform(rng, shape, 1.5)
    if "norm" in name or name.startswith("heatmap_conv.1."):
        if name.endswith("weight"):
            return (1.0 + _uniform(rng, shape, 0.25)).astype(np.float32)
        return _uniform(rng, shape, 0.2)
    if name.endswith("bias"):
        return _uniform(rng, shape, 0.2)
    return _uniform(rng, shape, float(np.sqrt(6.0 / (shape[-1] + shape[0]))))


def _default_init_tensor(name, shape, seed):
    """PyTorch's default initialisation of nn.Linear / nn.Conv2d / nn.GroupNorm (what the reference's tlbr_reg and
    heatmap_conv get: src/model.py:59-77 are outside the xavier loop of transformer.py:308-311): kaiming_uniform
    (a = sqrt(5)) -> bound 1/sqrt(fan_in) for weights and biases; GroupNorm weight 1, bias 0."""
    shape = tuple(shape)
    if name.startswith("heatmap_conv.1."):
        return np.ones(shape, np.float32) if name.endswith("weight") else np.zeros(shape, np.float32)
    fan_in = {"tlbr_reg.0": 256, "tlbr_reg.2": 256, "heatmap_conv.0": 256 * 9, "heatmap_conv.3": 256}[name.rsplit(".", 1)[0]]
    return _uniform(_rng_for(name + "/default", seed), shape, 1.0 / float(np.sqrt(fan_in)))


def synthetic_hot_path_weights(seed=0, include_unused=False, ln_gain=None, head_default_init=False):
    """{name: fp32 ndarray} for every hot-path tensor (random-init weights of the reference architecture).
    Stress variants (tests/golden STRESS_CASES): ln_gain = (lo, hi) multiplies every LayerNorm weight by a deterministic
    per-channel factor in [lo, hi] (trained-like gains); head_default_init uses PyTorch's default initialisation for
    tlbr_reg / heatmap_conv instead of the test-friendly scales of synthetic_tensor."""
    names = list(CANONICAL_ORDER) + (UNUSED_NAMES if include_unused else [])
    out = {n: synthetic_tensor(n, s, seed) for n, s in names}
    if ln_gain is not None:
        lo, hi = ln_gain
        for n, s in names:
            if "norm" in n and n.endswith("weight") and n.startswith("transformer."):
                u = _uniform(_rng_for(n + "/gain", seed), s, 1.0).astype(np.float64) * 0.5 + 0.5        # [0, 1)
                out[n] = (out[n] * (lo + (hi - lo) * u)).astype(np.float32)
    if head_default_init:
        for n, s in names:
            if n.startswith(("tlbr_reg.", "heatmap_conv.")):
                out[n] = _default_init_tensor(n, s, seed)
    return out


def synthetic_features(batch, hf, wf, seed=1, tag="feat", scale=1.0):
    """Stand-in for input_proj2 outputs: 0.3*N(0,1)-like (real features have std ~0.30); `scale`
    multiplies them (stress cases)."""
    rng = _rng_for("%s/%d/%d/%d" % (tag, batch, hf, wf), seed)
    shape = (batch, D_MODEL, hf, wf)
    # sum of 4 uniforms: bell-shaped, exactly reproducible
    acc = sum(_uniform(rng, shape, 1.0).astype(np.float64) for _ in range(4))
    return (acc * (scale * 0.3 / np.sqrt(4.0 / 3.0))).astype(np.float32)


def synthetic_mask(batch, hf, wf, tag="mask"):
    """Deterministic padding-style masks [batch, hf, wf] (float32 0/1): image b keeps the top-left
    (hf - r_b) x (wf - c_b) positions, with r_b, c_b small and depending on b and the tag."""
    off = sum(ord(ch) for ch in tag) % 3
    m = np.zeros((batch, hf, wf), dtype=np.float32)
    for b in range(batch):
        r = min(hf - 1, (b + 1 + off) % 4)
        c = min(wf - 1, (2 * b + 1 + off) % 5)
        m[b, : hf - r, : wf - c] = 1.0
    return m


# --------------------------------------------------------------------------------------------------------
# neck: input_proj -> PatchMerging -> input_proj2, reference src/model.py:45-56,
# src/models/backbone.py:28-51.  NECK_ORDER is the packing order of oetr_neck_create (include/oetr_b200.h).
# --------------------------------------------------------------------------------------------------------
BACKBONE_CH = 1024
NECK_ORDER = [
    ("input_proj.weight", (256, 1024, 1, 1)), ("input_proj.bias", (256,)),
    ("patchmerging.norm.weight", (256,)), ("patchmerging.norm.bias", (256,)),
    ("patchmerging.reductions.0.weight", (256, 256, 4, 4)), ("patchmerging.reductions.0.bias", (256,)),
    ("patchmerging.reductions.1.weight", (128, 256, 8, 8)), ("patchmerging.reductions.1.bias", (128,)),
    ("patchmerging.reductions.2.weight", (128, 256, 16, 16)), ("patchmerging.reductions.2.bias", (128,)),
    ("input_proj2.weight", (256, 512, 1, 1)), ("input_proj2.bias", (256,)),
]
NECK_PACKED_COUNT = sum(int(np.prod(s)) for _, s in NECK_ORDER)      # 11 929 088


def pack_neck_weights(state_dict):
    """state_dict (reference key names) -> one fp32 ndarray in NECK_ORDER."""
    parts = []
    for name, shape in NECK_ORDER:
        if name not in state_dict:
            raise KeyError("neck weight %r missing from state dict" % name)
        v = state_dict[name]
        if hasattr(v, "detach"):
            v = v.detach().to("cpu").float().numpy()
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(shape):
            raise ValueError("weight %r has shape %s, expected %s" % (name, tuple(v.shape), shape))
        parts.append(v.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def synthetic_neck_weights(seed=0, gain=1.0):
    """{name: fp32 ndarray}: PyTorch's default nn.Conv2d initialisation (bound 1/sqrt(fan_in) for weight and bias --
    what the reference's neck gets, it is outside the xavier loop) times `gain`; LayerNorm scale around 1."""
    out = {}
    for name, shape in NECK_ORDER:
        rng = _rng_for(name, seed)
        if ".norm." in name:
            out[name] = (1.0 + _uniform(rng, shape, 0.25)) if name.endswith("weight") else _uniform(rng, shape, 0.2)
            out[name] = out[name].astype(np.float32)
            continue
        wshape = dict(NECK_ORDER)[name.rsplit(".", 1)[0] + ".weight"]
        fan_in = int(np.prod(wshape[1:]))
        out[name] = (_uniform(rng, shape, 1.0 / float(np.sqrt(fan_in))) * np.float32(gain)).astype(np.float32)
    return out


def synthetic_backbone_features(batch, h, w, seed=1, tag="layer3", scale=1.0):
    """Stand-in for ResNet-50 layer3 outputs [batch,1024,h,w]: post-ReLU, about half the entries zero."""
    rng = _rng_for("%s/%d/%d/%d" % (tag, batch, h, w), seed)
    shape = (batch, BACKBONE_CH, h, w)
    acc = sum(_uniform(rng, shape, 1.0).astype(np.float64) for _ in range(3))
    return (np.maximum(acc, 0.0) * scale).astype(np.float32)
