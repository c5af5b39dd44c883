"""dloc overlap plugin backed by the B200 hot path: same conf keys, weights file convention and tuple return as
the reference plugin (reference dloc/core/overlaps/oetr.py:15-46), so `--overlaper oetr` callers
(evaluation.py:41-45,77-80; dloc/core/overlap_features.py:267-294) work unchanged."""
import torch

# Absolute imports only: this file works where it lies (oetr_b200.dloc.core.overlaps.oetr) AND copied / linked into the
# reference tree as dloc/core/overlaps/<name>.py, where the host's `dynamic_load` must find exactly one subclass of the
# HOST's BaseModel in it (reference dloc/core/utils/base_model.py:37-46).  `oetr_b200` must be importable (repo root on
# sys.path).
from oetr_b200.config import get_cfg_defaults
from oetr_b200.dloc.core.utils.base_model import BaseModel        # the host tree's class when `dloc` is importable
from oetr_b200.model import build_detectors


class OETR(BaseModel):
    default_conf = {
        'model': 'oetr',
        'num_layers': 50,
        'stride': 32,
        'last_layer': 1024,
        'weights': 'oetr.pth',
        # additions of this build (ignored by reference callers):
        'attention': 'linear',     # QueryTransformer attention_mode
        'precision': 'fp16',       # 'fp16' tcgen05 path | 'fp32' CUDA-core path
    }
    required_inputs = ['image0', 'image1']

    def build_cfg(self, conf):
        cfg = get_cfg_defaults()
        cfg.OETR.MODEL = conf['model']
        cfg.OETR.BACKBONE.STRIDE = conf['stride']
        cfg.OETR.BACKBONE.LAYER = conf['layer']      # like the reference, callers must supply 'layer'
        cfg.OETR.BACKBONE.LAST_LAYER = conf['last_layer']
        cfg.OETR.BACKBONE.NUM_LAYERS = conf['num_layers']
        return cfg

    def _init(self, conf, model_path):
        self.conf = {**self.default_conf, **conf}
        self.cfg = self.build_cfg(self.conf)
        self.net = build_detectors(self.cfg.OETR, attention_mode=self.conf['attention'],
                                   precision=self.conf['precision'])
        self.net.load_state_dict(torch.load(model_path / self.conf['weights'], map_location='cpu'))

    def _forward(self, data):
        return self.net.forward_dummy(data['image0'], data['image1'])
