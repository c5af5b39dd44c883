"""Importable name of the `imagematching-oetr_b200/` package directory (a hyphen is not a valid identifier):
this package's search path IS that directory, so `oetr_b200.model`, `oetr_b200.dloc...` are its modules."""
import os

__path__ = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imagematching-oetr_b200")]

from . import cabi, weights  # noqa: E402,F401
from .config import get_cfg_defaults  # noqa: E402,F401
from .hotpath import OverlapHotPath  # noqa: E402,F401
from .neck import NeckB200  # noqa: E402,F401
from .model import OETR, QueryTransformer, build_detectors  # noqa: E402,F401
