"""Host-side mirror of the reference interfaces: parameter names, packing, plugin convention, loud failure
without a GPU.  CPU only."""
import numpy as np
import pytest
import torch

import oetr_b200
from oetr_b200 import weights
from oetr_b200.dloc.core import overlap_features, overlaps
from oetr_b200.dloc.core.utils.base_model import BaseModel, dynamic_load

from conftest import ROOT as ROOT_DIR

try:
    import ref_loader
    HAVE_REF = ref_loader.reference_available()
except Exception:  # pragma: no cover
    HAVE_REF = False


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return oetr_b200.build_detectors(oetr_b200.get_cfg_defaults().OETR).eval()


def test_state_dict_has_every_hot_path_name(model):
    sd = model.state_dict()
    for name, shape in weights.CANONICAL_ORDER + weights.UNUSED_NAMES + weights.NECK_ORDER:
        assert name in sd and tuple(sd[name].shape) == tuple(shape), name
    assert "pos_encoding.pe" not in sd          # non-persistent in the reference (models/utils.py:196)


def test_pack_order_and_errors(model):
    sd = model.state_dict()
    packed = weights.pack_hot_path_weights(sd)
    assert packed.dtype == np.float32 and packed.size == weights.PACKED_COUNT
    first = sd["transformer.encoder.0.q_proj.weight"].numpy().reshape(-1)
    assert np.array_equal(packed[: first.size], first)
    assert packed[-1] == sd["heatmap_conv.3.bias"].item()
    bad = dict(sd)
    del bad["tlbr_reg.2.bias"]
    with pytest.raises(KeyError):
        weights.pack_hot_path_weights(bad)
    bad = dict(sd)
    bad["tlbr_reg.2.bias"] = torch.zeros(5)
    with pytest.raises(ValueError):
        weights.pack_hot_path_weights(bad)


def test_synthetic_weights_are_deterministic():
    a = weights.synthetic_hot_path_weights(3)
    b = weights.synthetic_hot_path_weights(3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert not np.array_equal(a["tlbr_reg.0.weight"], weights.synthetic_hot_path_weights(4)["tlbr_reg.0.weight"])


def test_unknown_model_raises_value_error():
    cfg = oetr_b200.get_cfg_defaults()
    cfg.OETR.MODEL = "oetr_fcos"
    with pytest.raises(ValueError):
        oetr_b200.build_detectors(cfg.OETR)            # reference src/model.py:384


def test_hot_path_fails_loudly_without_cuda(model):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    img = torch.rand(1, 64, 64, 3)
    with pytest.raises(RuntimeError):
        model.forward_dummy(img, img)
    with pytest.raises(RuntimeError):
        model.forward_dummy(img, img, mask1=torch.ones(1, 2, 2), mask2=torch.ones(1, 2, 2))


def test_plugin_convention():
    cls = dynamic_load(overlaps, "oetr")
    assert issubclass(cls, BaseModel) and cls.required_inputs == ["image0", "image1"]
    conf = overlap_features.confs["oetr"]["model"]
    assert {"name", "model", "stride", "last_layer", "num_layers", "layer", "weights"} <= set(conf)
    assert cls.default_conf["weights"] == "oetr.pth"


class _Needs(BaseModel):
    required_data_keys = ["image0"]

    def __init__(self):
        super().__init__({}, None)

    def _init(self, conf, model_path):
        pass

    def _forward(self, data):
        return data


def test_plugin_loads_checkpoint_from_model_path(tmp_path, model):
    (tmp_path / "oetr").mkdir()
    torch.save(model.state_dict(), tmp_path / "oetr" / "x.pth")
    conf = dict(overlap_features.confs["oetr"]["model"], weights="oetr/x.pth")
    plug = dynamic_load(overlaps, "oetr")(conf, tmp_path)
    a, b = plug.net.state_dict(), model.state_dict()
    assert all(torch.equal(a[k], b[k]) for k in b)
    with pytest.raises(AssertionError):
        _Needs()({})                                   # base_model.py:22-23 missing-key assert


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_strict_load_and_feature_extraction_match_reference(model):
    ref = ref_loader.build_reference_oetr(seed=0)
    model.load_state_dict(ref.state_dict(), strict=True)
    assert set(model.state_dict()) == set(ref.state_dict())
    torch.manual_seed(1)
    a, b = torch.rand(1, 192, 256, 3), torch.rand(1, 256, 192, 3)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.feature_extraction(a, b)                    # the default CUDA neck refuses to run on the CPU
        model.neck_mode = "torch"                             # the reference's PyTorch modules (bit-exact check below)
        m = model.feature_extraction(a, b)
        model.neck_mode = "cuda"
        r = ref.feature_extraction(a, b)
    assert len(m) == len(r) == 8                              # (feat1, feat2, pos1, pos2, hf1, wf1, hf2, wf2), model.py:130
    assert torch.equal(m[0], r[0]) and torch.equal(m[1], r[1])
    assert torch.allclose(m[2], r[2], atol=1e-6, rtol=0) and torch.allclose(m[3], r[3], atol=1e-6, rtol=0)
    assert tuple(m[4:]) == tuple(r[4:])
    assert torch.allclose(model.pos_encoding.pe, ref.pos_encoding.pe, atol=1e-6, rtol=0)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_plugin_is_found_by_the_reference_dynamic_load(tmp_path, model):
    """The plugin file, placed FIRST on the reference's `dloc.core.overlaps` package path (= copied over
    dloc/core/overlaps/oetr.py), is what the reference's own dynamic_load returns for `--overlaper oetr`: one subclass of
    the HOST's BaseModel, same default_conf keys and tuple-returning `_forward` (reference evaluation.py:41-45)."""
    import importlib
    import os
    import subprocess
    import sys
    code = """
import sys, os, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests', 'golden')); sys.path.insert(0, %r)
import ref_loader                                   # installs the kornia / timm / yacs shims the reference needs
sys.path.insert(0, '/root/reference')
import dloc.core.overlaps as host_overlaps
from dloc.core.utils.base_model import BaseModel as HostBase, dynamic_load as host_dynamic_load
import oetr_b200
plug_dir = os.path.join(os.path.dirname(oetr_b200.__path__[0]), 'imagematching-oetr_b200', 'dloc', 'core', 'overlaps')
host_overlaps.__path__.insert(0, plug_dir)
cls = host_dynamic_load(host_overlaps, 'oetr')
assert cls.__module__ == 'dloc.core.overlaps.oetr' and issubclass(cls, HostBase), cls.__mro__
assert os.path.samefile(sys.modules[cls.__module__].__file__, os.path.join(plug_dir, 'oetr.py'))
from oetr_b200.dloc.core import overlap_features
conf = dict(overlap_features.confs['oetr']['model'], weights='oetr/x.pth')
plug = cls(conf, __import__('pathlib').Path(%r))
assert isinstance(plug.net, oetr_b200.OETR) and {'model', 'num_layers', 'stride', 'last_layer', 'weights'} <= set(cls.default_conf)
print('OK')
""" % (ROOT_DIR, ROOT_DIR, ROOT_DIR, str(tmp_path))
    (tmp_path / "oetr").mkdir()
    torch.save(model.state_dict(), tmp_path / "oetr" / "x.pth")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stderr[-3000:]
