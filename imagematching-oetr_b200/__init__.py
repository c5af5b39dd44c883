"""oetr-b200: B200-native OETR hot path (pair-wise feature-correlation transformer + overlap-box head) behind
the reference's Python interfaces.  The directory name contains a hyphen; import it as `oetr_b200` (alias
package at the repo root) or via importlib.import_module("imagematching-oetr_b200")."""
from . import cabi, weights  # noqa: F401
from .config import get_cfg_defaults  # noqa: F401
from .hotpath import OverlapHotPath  # noqa: F401
from .neck import NeckB200  # noqa: F401
from .model import OETR, QueryTransformer, build_detectors  # noqa: F401
