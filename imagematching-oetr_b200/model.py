"""Drop-in mirror of the reference model class (reference src/model.py:38-384): same constructor argument,
same parameter names (strict load_state_dict of reference checkpoints), same `forward_dummy` / `forward`
signatures and return types.  feature_extraction stays PyTorch; everything after it (feature_correlation,
center_estimation, size_regression, box assembly) is one call into the CUDA library."""
import torch
import torch.nn as nn

from .backbone import PatchMerging, ResnetEncoder
from .hotpath import OverlapHotPath
from .weights import CANONICAL_ORDER, UNUSED_NAMES


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


class QueryTransformer(nn.Module):
    """Weights of the reference QueryTransformer (transformer.py:287-311).  It holds parameters only: the
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

    def forward(self, *args, **kwargs):
        raise RuntimeError("QueryTransformer here is a parameter container; call OETR.forward_dummy / "
                           "OETR.feature_correlation_and_regression (CUDA hot path)")


class OETR(nn.Module):
    def __init__(self, cfg, attention_mode="linear", precision="fp16", pretrained_backbone=False):
        super().__init__()
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
        self.cycle = cfg.LOSS.CYCLE_OVERLAP
        self.softmax_temperature = 1
        self.precision = precision
        self._hot = None

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
        return self._hot

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._hot = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._hot = None
        return out

    @property
    def hot_path(self):
        return self._hot if self._hot is not None else self.refresh_hot_path()

    # ---- stages ---------------------------------------------------------------------------------------
    def feature_extraction(self, image1, image2, mask1=None, mask2=None):
        """backbone -> input_proj -> patchmerging -> input_proj2 for both images (model.py:109-130)."""
        feats = []
        for img in (image1, image2):
            f = self.input_proj2(self.patchmerging(self.input_proj(self.backbone(img))))
            feats.append(f)
        return feats[0], feats[1]

    def feature_correlation_and_regression(self, feat1, feat2, hw1, hw2, clamp=True, debug=False, mask1=None,
                                           mask2=None):
        return self.hot_path.forward(feat1, feat2, hw1, hw2, clamp=clamp, debug=debug, mask1=mask1, mask2=mask2)

    @torch.no_grad()
    def forward_dummy(self, image1, image2, mask1=None, mask2=None):
        """Inference entry (model.py:229-252): NHWC fp32 images in [0,1] -> (box1, box2) clamped xyxy.
        mask1 / mask2 [B,hf,wf] (feature-map resolution, float; both or neither): the masks the reference hands to
        feature_correlation and center_estimation (model.py:240-247); linear attention only."""
        hw1, hw2 = tuple(image1.shape[1:3]), tuple(image2.shape[1:3])
        feat1, feat2 = self.feature_extraction(image1, image2)
        return self.feature_correlation_and_regression(feat1, feat2, hw1, hw2, clamp=True, mask1=mask1, mask2=mask2)

    @torch.no_grad()
    def forward(self, data, validation=False):
        """Training-signature entry (model.py:255-376) restricted to its inference half: unclamped boxes
        (model.py:193-211) and, when ground truth is given, the IoU metrics.  Losses are training-only and
        out of scope."""
        valid = data["overlap_valid"] if "overlap_valid" in data else slice(None)
        image1, image2 = data["image1"][valid], data["image2"][valid]
        if "resize_mask1" in data:
            raise NotImplementedError("resize_mask inputs are never produced by the reference datasets")
        hw1, hw2 = tuple(image1.shape[1:3]), tuple(image2.shape[1:3])
        feat1, feat2 = self.feature_extraction(image1, image2)
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
