"""Drop-in mirror of the reference model class (reference src/model.py:38-384): same constructor argument,
same parameter names (strict load_state_dict of reference checkpoints), same `forward_dummy` / `forward`
signatures and return types.  The ResNet trunk stays PyTorch; the neck (input_proj -> patchmerging -> input_proj2) is one
call into the CUDA library (oetr_neck_forward) and everything after it (feature_correlation, center_estimation,
size_regression, box assembly) another (oetr_forward)."""
import math

import torch
import torch.nn as nn

from .backbone import PatchMerging, ResnetEncoder
from .hotpath import OverlapHotPath
from .neck import NeckB200
from .weights import CANONICAL_ORDER, NECK_ORDER, UNUSED_NAMES


def _linear(i, o, bias):
    return nn.Linear(i, o, bias=bias)


class _EncoderLayerParams(nn.Module):
    """Parameter container with the names of reference EncoderLayer (transformer.py:75-102)."""

    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.merge = (_linear(d, d, False) for _ in range(4))
        self.mlp = nn.Sequential(_linear(d, 2 * d, False), nn.GELU(), _linear(2 * d, d, False))
        self.pre_norm_q, self.pre_norm_kv, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)


class _MHAParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj = (_linear(d, d, True) for _ in range(3))
        self.merge = _linear(d, d, False)


class _DecoderLayerParams(nn.Module):
    """Names of reference DecoderLayer (transformer.py:189-222), including its four never-used projections."""

    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj = (_linear(d, d, False) for _ in range(3))
        self.self_attn, self.multihead_attn = _MHAParams(d), _MHAParams(d)
        self.merge = _linear(d, d, False)
        self.mlp = nn.Sequential(_linear(d, 2 * d, False), nn.ReLU(True), _linear(2 * d, d, False))
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)


class _DecoderParams(nn.Module):
    def __init__(self, d, n):
        super().__init__()
        self.layers = nn.ModuleList(_DecoderLayerParams(d) for _ in range(n))


class PositionEncodingSine(nn.Module):
    """The constant table of reference src/models/utils.py:174-205, restated from its definition (SURVEY.md Appendix
    A.1): frequencies exp(-2k) (the reference's `-log(1e4)/d_model // 2` evaluates to -1), positions from 1, channels
    4k..4k+3 = sin(x d), cos(x d), sin(y d), cos(y d).  Non-persistent buffer (absent from checkpoints).  The CUDA path
    has its own copy of this table; this module exists so that feature_extraction returns the reference's 8-tuple."""

    def __init__(self, d_model, max_shape=(100, 100)):
        super().__init__()
        h, w = max_shape
        ys = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
        xs = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
        factor = (-math.log(10000.0) / d_model // 2)                     # == -1.0
        div = torch.exp(torch.arange(0, d_model // 2, 2).float() * factor)[:, None, None]
        pe = torch.zeros(d_model, h, w)
        pe[0::4], pe[1::4] = torch.sin(xs * div), torch.cos(xs * div)
        pe[2::4], pe[3::4] = torch.sin(ys * div), torch.cos(ys * div)
        self.register_buffer("pe", pe[None], persistent=False)

    def forward(self, x):
        return self.pe[:, :, :x.size(2), :x.size(3)]


def _versions(params):
    return tuple((p.data_ptr(), p._version) for p in params)


class QueryTransformer(nn.Module):
    """Weights of the reference QueryTransformer (transformer.py:287-311) with its forward signature; the
    arithmetic lives in liboetr_b200.so.  `attention_mode` selects linear (shipped default) or full attention."""

    def __init__(self, d_model=256, nhead=8, num_layers=4, attention_mode="linear"):
        super().__init__()
        assert d_model == 256 and nhead == 8 and num_layers == 4, "the CUDA path is specialised for OETR's sizes"
        self.attention_mode = attention_mode
        self.encoder = nn.ModuleList(_EncoderLayerParams(d_model) for _ in range(2 * num_layers))
        self.decoder = _DecoderParams(d_model, 2)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

        self._hot, self._hot_key = None, None
        self.precision = "fp16"

    def forward(self, feat0, feat1, query_embed0, query_embed1, pos0=None, pos1=None, mask0=None, mask1=None):
        """reference transformer.py:313-383: feat* [N,C,h,w], query_embed* [1,C] -> (hs0, hs1 [N,1,C], memory0, memory1
        [N,L,C]).  pos0 / pos1 are accepted for signature compatibility; the CUDA path uses its own
        PositionEncodingSine table (the only encoding the reference ever passes, src/model.py:126-127).
        Stand-alone use builds a CUDA handle from this module's weights and the given query embeddings (head weights
        zero) and rebuilds it when any of them changes; inside OETR the model's own handle is used instead."""
        from .weights import CANONICAL_ORDER
        params = list(self.parameters()) + [query_embed0, query_embed1]
        key = _versions(params)
        if self._hot is None or self._hot_key != key:
            if self._hot is not None:
                self._hot.close()
            sd = {"transformer." + k: v for k, v in self.state_dict().items()}
            sd["query_embed1.weight"], sd["query_embed2.weight"] = query_embed0.detach(), query_embed1.detach()
            for name, shape in CANONICAL_ORDER:
                if name not in sd:
                    sd[name] = torch.zeros(shape)
            self._hot = OverlapHotPath(sd, attention=self.attention_mode, precision=self.precision, device=feat0.device)
            self._hot_key = key
        hw0, hw1 = (32 * feat0.shape[2], 32 * feat0.shape[3]), (32 * feat1.shape[2], 32 * feat1.shape[3])
        _, _, dbg = self._hot.forward(feat0, feat1, hw0, hw1, clamp=False, debug=True, mask1=mask0, mask2=mask1)
        return dbg["hs1"][:, None], dbg["hs2"][:, None], dbg["memory1"], dbg["memory2"]


class OETR(nn.Module):
    def __init__(self, cfg, attention_mode="linear", precision="fp16", pretrained_backbone=False, neck="cuda"):
        """neck: "cuda" runs input_proj -> patchmerging -> input_proj2 in liboetr_b200.so (oetr_neck_forward; ResNet-50
        layer3 features only, no fallback), "torch" keeps the reference's PyTorch modules (needed for training the neck
        or for backbones with another channel count)."""
        super().__init__()
        if neck not in ("cuda", "torch"):
            raise ValueError("neck %r not in ('cuda', 'torch')" % (neck,))
        self.neck_mode = neck
        self._neck, self._neck_key = None, None
        self.backbone = ResnetEncoder(cfg, pretrained=pretrained_backbone)
        self.d_model = self.backbone.last_layer // 4
        self.input_proj = nn.Conv2d(self.backbone.last_layer, self.d_model, kernel_size=1)
        self.input_proj2 = nn.Conv2d(self.d_model * 2, self.d_model, kernel_size=1)
        self.patchmerging = PatchMerging((20, 20), self.d_model, norm_layer=nn.LayerNorm, patch_size=[4, 8, 16])
        self.tlbr_reg = nn.Sequential(nn.Linear(self.d_model, self.d_model, False), nn.ReLU(inplace=True),
                                      nn.Linear(self.d_model, 4))
        self.heatmap_conv = nn.Sequential(
            nn.Conv2d(self.d_model, self.d_model, (3, 3), padding=(1, 1), stride=(1, 1), bias=True),
            nn.GroupNorm(32, self.d_model), nn.ReLU(inplace=True), nn.Conv2d(self.d_model, 1, (1, 1)))
        self.query_embed1 = nn.Embedding(1, self.d_model)
        self.query_embed2 = nn.Embedding(1, self.d_model)
        self.transformer = QueryTransformer(self.d_model, nhead=8, num_layers=4, attention_mode=attention_mode)
        self.max_shape = tuple(cfg.NECK.MAX_SHAPE)
        self.pos_encoding = PositionEncodingSine(self.d_model, max_shape=self.max_shape)
        self.cycle = cfg.LOSS.CYCLE_OVERLAP
        self.softmax_temperature = 1
        self.precision = precision
        self._hot, self._hot_key = None, None
        self.h1 = self.w1 = self.h2 = self.w2 = None         # set by forward_dummy / forward like the reference (model.py:232-233)

    # ---- hot-path handle management ---------------------------------------------------------------------
    def refresh_hot_path(self):
        """(Re)build the CUDA handle from the current parameters.  Called lazily; call it yourself after
        mutating weights in place."""
        if self._hot is not None:
            self._hot.close()
        sd = {k: v for k, v in self.state_dict().items()}
        dev = self.query_embed1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("OETR hot path needs the module on a CUDA device (there is no CPU fallback); "
                               "call .cuda() first")
        self._hot = OverlapHotPath(sd, attention=self.transformer.attention_mode, precision=self.precision,
                                   max_shape=self.max_shape, device=dev)
        self._hot_key = _versions(self._hot_params())
        return self._hot

    def _hot_params(self):
        if getattr(self, "_hot_param_list", None) is None:
            names = {n for n, _ in CANONICAL_ORDER}
            self._hot_param_list = [p for n, p in self.named_parameters() if n in names]
        return self._hot_param_list

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._hot = self._neck = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._hot = self._neck = None
        return out

    def _neck_params(self):
        if getattr(self, "_neck_param_list", None) is None:
            names = {n for n, _ in NECK_ORDER}
            self._neck_param_list = [p for n, p in self.named_parameters() if n in names]
        return self._neck_param_list

    @property
    def neck_path(self):
        """The CUDA neck handle (a snapshot of input_proj / patchmerging / input_proj2), rebuilt like `hot_path` when one
        of those parameters changed."""
        if self._neck is None or self._neck_key != _versions(self._neck_params()):
            if self._neck is not None:
                self._neck.close()
            dev = self.input_proj.weight.device
            if dev.type != "cuda":
                raise RuntimeError("the CUDA neck needs the module on a CUDA device (there is no CPU fallback); call .cuda() "
                                   "or build the model with neck='torch'")
            if self.backbone.last_layer != 1024:
                raise RuntimeError("the CUDA neck is specialised for 1024-channel backbone features (ResNet-50 layer3), got "
                                   "%d; build the model with neck='torch'" % self.backbone.last_layer)
            self._neck = NeckB200({n: p for n, p in self.state_dict().items()}, device=dev)
            self._neck_key = _versions(self._neck_params())
        return self._neck

    @property
    def hot_path(self):
        """The CUDA handle snapshots the weights: it is rebuilt when a hot-path parameter was replaced or modified in
        place (optimizer.step, param.data.copy_) since the snapshot -- a cheap (data_ptr, _version) check per call."""
        if self._hot is None or self._hot_key != _versions(self._hot_params()):
            self.refresh_hot_path()
        return self._hot

    # ---- stages ---------------------------------------------------------------------------------------
    def feature_extraction(self, image1, image2, mask1=None, mask2=None):
        """backbone -> input_proj -> patchmerging -> input_proj2 for both images and the position encodings: the
        reference's 8-tuple (feat1, feat2, pos1, pos2, hf1, wf1, hf2, wf2), model.py:109-130."""
        if self.neck_mode == "cuda" and not (torch.is_grad_enabled() and self.training):
            if image1.shape == image2.shape:
                # both image sets through the trunk and the neck as ONE batch (eval-mode BatchNorm: samples are independent);
                # concatenating the images (157 MB at batch 32) is cheap, concatenating the 1024-channel features is not
                n = image1.shape[0]
                f = self.neck_path.forward(self.backbone(torch.cat([image1, image2], dim=0)))
                feat1, feat2 = f[:n], f[n:]
            else:
                feat1, feat2 = self.neck_path.forward(self.backbone(image1)), self.neck_path.forward(self.backbone(image2))
        else:
            feat1 = self.input_proj2(self.patchmerging(self.input_proj(self.backbone(image1))))
            feat2 = self.input_proj2(self.patchmerging(self.input_proj(self.backbone(image2))))
        hf1, wf1 = feat1.shape[2:]
        hf2, wf2 = feat2.shape[2:]
        return feat1, feat2, self.pos_encoding(feat1), self.pos_encoding(feat2), hf1, wf1, hf2, wf2

    @torch.no_grad()
    def feature_correlation(self, feat1, feat2, pos1=None, pos2=None, mask1=None, mask2=None):
        """model.py:132-143 -> (hs1, hs2 [N,1,256], memory1, memory2 [N,L,256]) from the CUDA path's stage taps.  pos1 /
        pos2 are the PositionEncodingSine slices the reference passes; the kernels read their own copy of that table."""
        hw1 = (self.h1 or 32 * feat1.shape[2], self.w1 or 32 * feat1.shape[3])
        hw2 = (self.h2 or 32 * feat2.shape[2], self.w2 or 32 * feat2.shape[3])
        _, _, dbg = self.hot_path.forward(feat1, feat2, hw1, hw2, clamp=False, debug=True, mask1=mask1, mask2=mask2)
        return dbg["hs1"][:, None], dbg["hs2"][:, None], dbg["memory1"], dbg["memory2"]

    def _head(self, hs1, hs2, memory1, memory2, hf1, wf1, hf2, wf2, mask1, mask2):
        if self.h1 is None:
            raise RuntimeError("set self.h1, w1, h2, w2 (image sizes) first, as forward_dummy does (model.py:232-233)")
        return self.hot_path.head(hs1, hs2, memory1, memory2, hf1, wf1, hf2, wf2, (self.h1, self.w1), (self.h2, self.w2),
                                  clamp=False, mask1=mask1, mask2=mask2)

    @torch.no_grad()
    def center_estimation(self, hs1, hs2, memory1, memory2, hf1, wf1, hf2, wf2, mask1=None, mask2=None):
        """model.py:145-186 -> (box_cxy1, box_cxy2) [N,2] pixels; uses self.h1 / self.h2 for the grid stride."""
        out = self._head(hs1, hs2, memory1, memory2, hf1, wf1, hf2, wf2, mask1, mask2)
        return out[2], out[3]

    @torch.no_grad()
    def size_regression(self, hs1, hs2):
        """model.py:188-191 -> (tlbr1, tlbr2) [N,4] in (0,1)."""
        n = hs1.shape[0]
        z = torch.zeros(n, 1, 256, device=hs1.device)
        h1, w1, h2, w2 = self.h1, self.w1, self.h2, self.w2
        if h1 is None:
            self.h1 = self.w1 = self.h2 = self.w2 = 32
        try:
            out = self._head(hs1, hs2, z, z, 1, 1, 1, 1, None, None)
        finally:
            self.h1, self.w1, self.h2, self.w2 = h1, w1, h2, w2
        return out[4], out[5]

    def feature_correlation_and_regression(self, feat1, feat2, hw1, hw2, clamp=True, debug=False, mask1=None,
                                           mask2=None):
        return self.hot_path.forward(feat1, feat2, hw1, hw2, clamp=clamp, debug=debug, mask1=mask1, mask2=mask2)

    @torch.no_grad()
    def forward_dummy(self, image1, image2, mask1=None, mask2=None):
        """Inference entry (model.py:229-252): NHWC fp32 images in [0,1] -> (box1, box2) clamped xyxy.
        mask1 / mask2 [B,hf,wf] (feature-map resolution, float; both or neither): the masks the reference hands to
        feature_correlation and center_estimation (model.py:240-247); linear attention only."""
        hw1, hw2 = tuple(image1.shape[1:3]), tuple(image2.shape[1:3])
        self.h1, self.w1 = hw1
        self.h2, self.w2 = hw2
        feat1, feat2 = self.feature_extraction(image1, image2)[:2]
        return self.feature_correlation_and_regression(feat1, feat2, hw1, hw2, clamp=True, mask1=mask1, mask2=mask2)

    @torch.no_grad()
    def forward(self, data, validation=False):
        """Training-signature entry (model.py:255-376) restricted to its inference half: unclamped boxes
        (model.py:193-211) and, when ground truth is given, the IoU metrics.  Losses are training-only and
        out of scope."""
        if self.training:
            raise NotImplementedError("oetr_b200.OETR is an inference drop-in: the CUDA hot path has no backward pass and the "
                                      "losses of reference src/model.py:300-376 are out of scope; call .eval()")
        valid = data["overlap_valid"] if "overlap_valid" in data else slice(None)
        image1, image2 = data["image1"][valid], data["image2"][valid]
        if "resize_mask1" in data:
            raise NotImplementedError("resize_mask inputs are never produced by the reference datasets")
        hw1, hw2 = tuple(image1.shape[1:3]), tuple(image2.shape[1:3])
        self.h1, self.w1 = hw1
        self.h2, self.w2 = hw2
        feat1, feat2 = self.feature_extraction(image1, image2)[:2]
        box1, box2 = self.feature_correlation_and_regression(feat1, feat2, hw1, hw2, clamp=False)
        out = {"pred_bbox1": box1, "pred_bbox2": box2}
        if "overlap_box1" in data:
            out["iou1"] = _aligned_iou(box1, data["overlap_box1"][valid].to(box1)).mean()
            out["iou2"] = _aligned_iou(box2, data["overlap_box2"][valid].to(box2)).mean()
        return out


def _aligned_iou(a, b, eps=1e-6):
    lt = torch.max(a[:, :2], b[:, :2])
    rb = torch.min(a[:, 2:], b[:, 2:])
    inter = (rb - lt).clamp(min=0).prod(dim=1)
    area = lambda x: (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])
    return inter / (area(a) + area(b) - inter).clamp(min=eps)


def build_detectors(cfg, **kwargs):
    """model.py:380-384"""
    if cfg.MODEL == "oetr":
        return OETR(cfg, **kwargs)
    raise ValueError(f"OETR.MODEL {cfg.MODEL} not supported.")


def hot_path_state_names():
    return [n for n, _ in CANONICAL_ORDER] + [n for n, _ in UNUSED_NAMES]
