"""``gsplat.rendering`` surface used by the reference: just ``rasterization``."""
import sys
from pathlib import Path

_repo = Path(__file__).resolve().parents[2]
if str(_repo) not in sys.path:
    sys.path.insert(0, str(_repo))

from easy_gaussian_splatting_b200.rendering import rasterization  # noqa: E402,F401

__all__ = ["rasterization"]
