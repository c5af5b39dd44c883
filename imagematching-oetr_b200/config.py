"""Configuration tree consumed by the OETR mirror (same keys as the reference's yacs defaults,
reference src/config/default.py:3-69; only the OETR subtree is needed on the inference path)."""
import copy


class CfgNode(dict):
    """Attribute-style nested dict; uses yacs when present, this otherwise (yacs is not in the image)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def clone(self):
        return copy.deepcopy(self)


def _node(**kw):
    n = CfgNode()
    n.update(kw)
    return n


def get_cfg_defaults():
    cfg = _node(OUTPUT="")
    cfg.OETR = _node(
        CHECKPOINT=None, BACKBONE_TYPE="ResNet", MODEL="oetr", NORM_INPUT=True,
        BACKBONE=_node(NUM_LAYERS=50, STRIDE=16, LAYER="layer3", LAST_LAYER=1024),
        NECK=_node(D_MODEL=256, LAYER_NAMES=["self", "cross"] * 4, ATTENTION="linear", MAX_SHAPE=(100, 100)),
        HEAD=_node(D_MODEL=256, NORM_REG_TARGETS=True),
        LOSS=_node(OIOU=False, CYCLE_OVERLAP=False, FOCAL_ALPHA=0.25, FOCAL_GAMMA=2.0, REG_WEIGHT=1.0,
                   CENTERNESS_WEIGHT=1.0),
    )
    return cfg
