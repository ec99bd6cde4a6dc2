"""vent_analysis_b200 -- B200-native voxel hot path of thomenr/Vent_Analysis
(N4 bias correction -> VDP / linear-binning VDP -> cluster index) behind the reference's own
class surface.  ``from vent_analysis_b200 import Vent_Analysis`` mirrors the reference's
``__init__.py:1``; ``vent_analysis_b200.CI`` mirrors its ``CI`` module."""
from .Vent_Analysis import Vent_Analysis  # noqa: F401
from . import CI  # noqa: F401

__all__ = ["Vent_Analysis", "CI"]
