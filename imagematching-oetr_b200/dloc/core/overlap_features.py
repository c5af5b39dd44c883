"""Named overlap-estimator configurations (reference dloc/core/overlap_features.py:20-46) and a batched
box-extraction helper.  The reference's per-pair crop / re-match pipeline (process, :49-261) is host-side image
plumbing and stays with the caller (out of scope)."""
import torch

from . import overlaps
from .utils.base_model import dynamic_load

confs = {
    'oetr_imc': {
        'output': 'oetr',
        'model': {'name': 'oetr', 'model': 'oetr', 'stride': 32, 'last_layer': 1024, 'num_layers': 50,
                  'layer': 'layer3', 'weights': 'oetr/sacdetrnet_mf_epoch24_2x4_best.pth'},
    },
    'oetr': {
        'output': 'oetr',
        'model': {'name': 'oetr', 'model': 'oetr', 'stride': 32, 'last_layer': 1024, 'num_layers': 50,
                  'layer': 'layer3', 'weights': 'oetr/sacdetrnet_mf_epoch30_2x4_cyclecenter.pth'},
    },
}


def build_overlap_model(conf, model_path, device='cuda'):
    """What evaluation.py:41-45 does for the overlap estimator."""
    Model = dynamic_load(overlaps, conf['model']['name'])
    return Model(conf['model'], model_path).eval().to(device)


@torch.no_grad()
def estimate_overlap(model, image0, image1):
    """(bbox0, bbox1) for NHWC [B,H,W,3] images in [0,1] -- the call at evaluation.py:77-80."""
    return model({'image0': image0, 'image1': image1})
