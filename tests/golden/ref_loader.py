"""Loader for the UNMODIFIED reference (read-only at /root/reference) -- test infrastructure only.

Installs the three sys.modules shims the reference needs in this container (kornia, timm, yacs are not
installed) and patches torchvision's resnet50 so no weight download is attempted
(reference src/models/backbone.py:145 passes pretrained=True).  Only used by tests/golden/make_golden.py and by
CPU tests that are skipped when /root/reference is absent (it does not exist on the GPU box).
"""
import copy
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("OETR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _meshgrid(height, width, normalized_coordinates=True, device="cpu", dtype=torch.float32):
    # kornia.utils.create_meshgrid: [1,H,W,2], last dim (x,y)
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    return torch.stack(torch.meshgrid([xs, ys], indexing="ij"), -1).permute(1, 0, 2)[None]


class CfgNode(dict):
    """Minimal yacs.config.CfgNode stand-in (attribute access + clone)."""

    def __getattr__(self, k):
        if k in self:
            return self[k]
        raise AttributeError(k)

    __setattr__ = dict.__setitem__

    def clone(self):
        return copy.deepcopy(self)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torchvision.models as tvm
    if "kornia" not in sys.modules:
        _mod("kornia", utils=_mod("kornia.utils", create_meshgrid=_meshgrid))
    if "timm" not in sys.modules:
        _mod("timm")
        _mod("timm.models")
        _mod("timm.models.layers", to_2tuple=lambda x: x if isinstance(x, (tuple, list)) else (x, x))
    if "yacs" not in sys.modules:
        _mod("yacs", config=_mod("yacs.config", CfgNode=CfgNode))
    r50 = tvm.resnet50
    if not getattr(r50, "_oetr_patched", False):
        def _resnet50(*a, **k):
            return r50(weights=None)
        _resnet50._oetr_patched = True
        tvm.resnet50 = _resnet50
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def build_reference_oetr(seed=0):
    install()
    from src.config.default import get_cfg_defaults
    from src.model import build_detectors
    cfg = get_cfg_defaults()
    cfg.OETR.BACKBONE.STRIDE = 32
    torch.manual_seed(seed)
    net = build_detectors(cfg.OETR).eval()
    return net


def build_reference_full_transformer(net):
    """QueryTransformer(attention_mode='full') carrying the same weights."""
    install()
    from src.models.transformer import QueryTransformer
    full = QueryTransformer(256, 8, 4, attention_mode="full")
    full.load_state_dict(net.transformer.state_dict())
    return full.eval()


def install_dloc():
    """Additionally stub matplotlib (not installed here; reference dloc/core/utils/utils.py:53-58 imports it for its
    plotting helpers only) so that the reference's dloc.core.utils.utils can be imported."""
    install()
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _mod("matplotlib", use=lambda *a, **k: None, pyplot=_mod("matplotlib.pyplot"), cm=_mod("matplotlib.cm"))
